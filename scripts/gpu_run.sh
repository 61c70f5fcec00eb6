mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
CROG_NO_FORK=1 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v4_raw.csv python tests/prof_forward.py 64 gpurun_out/ops_v4.tsv > gpurun_out/prof_fwd.log 2>&1
tail -1 gpurun_out/prof_fwd.log
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_tc -c 1 -o gpurun_out/gemm_l1c3_v4 -f python tests/prof_gemm_shape.py 692224 256 64 0 > gpurun_out/ncu_g.log 2>&1
tail -1 gpurun_out/ncu_g.log
