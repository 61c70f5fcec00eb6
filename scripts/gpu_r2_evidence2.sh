# Round-2 evidence, final kernels: launch lists + full captures of what changed after the first pass.
mkdir -p gpurun_out/ev
NCU="ncu --clock-control none --profile-from-start off"
FULL="$NCU --set full --import-source on -f"
CROG_NO_FORK=1 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --csv --log-file gpurun_out/ev/fwd_raw.csv python tests/prof_forward.py 64 gpurun_out/ev/fwd_ops.tsv > gpurun_out/ev/fwd.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/ev/bench_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --ncu-range > gpurun_out/ev/bench.log 2>&1
$FULL -k regex:warp_cubic -c 1 -o gpurun_out/ev/warp python tests/prof_kernels.py warp > /dev/null 2>&1
$FULL -k regex:preprocess_u8 -c 1 -o gpurun_out/ev/preprocess python tests/prof_kernels.py preprocess > /dev/null 2>&1
$FULL -k regex:gemm_tc -c 1 -o gpurun_out/ev/gemm_vis3_band python tests/prof_conv3.py 8 104 104 512 256 11 > /dev/null 2>&1
$FULL -k regex:gemm_tc -c 1 -o gpurun_out/ev/gemm_vis3_pertap python tests/prof_conv3.py 8 104 104 512 256 4 > /dev/null 2>&1
ls -la gpurun_out/ev | tail -8
