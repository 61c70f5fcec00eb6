# Snake-order experiment: DRAM bytes per GEMM launch with the L2 left as the previous kernel left it (--cache-control none).
mkdir -p gpurun_out/snake
NCU="ncu --clock-control none --cache-control none --profile-from-start off"
for s in 0 1; do
CROG_SNAKE=$s $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/snake/raw$s.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --ncu-range > gpurun_out/snake/bench$s.log 2>&1
done
ls -la gpurun_out/snake
