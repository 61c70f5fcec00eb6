mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
CROG_NO_FORK=1 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v3_raw.csv python tests/prof_forward.py 64 gpurun_out/ops_v3.tsv > gpurun_out/prof_fwd.log 2>&1
tail -1 gpurun_out/prof_fwd.log
ncu --metrics $M,smsp__inst_executed.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/tail_v3_raw.csv python tests/prof_tail.py 4096 > gpurun_out/prof_tail.log 2>&1
tail -2 gpurun_out/prof_tail.log
ncu --set full --import-source on --clock-control none -k regex:attention_tc -c 1 -o gpurun_out/attn_final -f python tests/prof_attn.py > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
