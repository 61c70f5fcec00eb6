for t in "" j64 j64b j256; do
so=""; [ -n "$t" ] && so=$PWD/crog_b200/lib/libcrog_b200.$t.so
echo "variant [$t]"; CROG_B200_SO=$so python scripts/time_jac.py
done
CROG_B200_SO=$PWD/crog_b200/lib/libcrog_b200.j64.so python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tail_golden.py -m gpu -q -k "jaccard or iou or golden" 2>&1 | tail -2
