mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
CROG_NO_FORK=1 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v5_raw.csv python tests/prof_forward.py 64 gpurun_out/ops_v5.tsv > gpurun_out/prof_fwd.log 2>&1
tail -1 gpurun_out/prof_fwd.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/bench_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --ncu-range > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
wc -l gpurun_out/bench_launches_raw.csv
