# evict_first hints A/B (same box): bench forward time, then DRAM bytes with the L2 left warm (--cache-control none)
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tile_cfgs or fused_downsample or two_operand" 2>&1 | tail -2
NH=$PWD/crog_b200/lib/libcrog_b200.nohint.so
for i in 1 2; do
for so in "$NH" ""; do
CROG_B200_SO=$so python bench.py --steps 40 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(\"hints\", int(\"$so\"==\"\"), d[\"ms_per_step\"], d[\"forward_ms_per_step\"], d[\"roofline\"][\"frac\"])"
done; done
mkdir -p gpurun_out/snake
NCU="ncu --clock-control none --cache-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/snake/raw2.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --ncu-range > gpurun_out/snake/bench2.log 2>&1
