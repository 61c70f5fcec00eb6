# evict_last store hints A/B (same box)
for i in 1 2; do
for t in "" el1 el0; do
so=""; [ -n "$t" ] && so=$PWD/crog_b200/lib/libcrog_b200.$t.so
CROG_B200_SO=$so python bench.py --steps 40 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(\"variant [$t]\", d[\"ms_per_step\"], d[\"forward_ms_per_step\"], d[\"roofline\"][\"frac\"])"
done; done
mkdir -p gpurun_out/snake
NCU="ncu --clock-control none --cache-control none --profile-from-start off"
CROG_B200_SO=$PWD/crog_b200/lib/libcrog_b200.el1.so $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/snake/raw3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --ncu-range > gpurun_out/snake/bench3.log 2>&1
