# Round-2 evidence pass: launch lists and `ncu --set full` captures (one GPU; nothing printed here is a bench value).
mkdir -p gpurun_out/ev
NCU="ncu --clock-control none --profile-from-start off"
FULL="$NCU --set full --import-source on -f"
# 1. every launch of one eager forward (plan order) with DRAM bytes and tensor-pipe activity
CROG_NO_FORK=1 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed --csv --log-file gpurun_out/ev/fwd_raw.csv python tests/prof_forward.py 64 gpurun_out/ev/fwd_ops.tsv > gpurun_out/ev/fwd.log 2>&1
# 2. launch list of bench.py's own timed steps (graph replay)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/ev/bench_raw.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --ncu-range > gpurun_out/ev/bench.log 2>&1
# 3. full captures of the non-GEMM kernels at bench shapes
$FULL -k regex:sigmoid_bicubic -c 1 -o gpurun_out/ev/glue python tests/prof_kernels.py glue > /dev/null 2>&1
$FULL -k regex:layernorm_chain -c 1 -o gpurun_out/ev/lnchain python tests/prof_kernels.py lnchain > /dev/null 2>&1
$FULL -k regex:gaussian -c 2 -o gpurun_out/ev/gaussian python tests/prof_kernels.py gaussian > /dev/null 2>&1
$FULL -k regex:warp_cubic -c 1 -o gpurun_out/ev/warp python tests/prof_kernels.py warp > /dev/null 2>&1
$FULL -k regex:preprocess_u8 -c 1 -o gpurun_out/ev/preprocess python tests/prof_kernels.py preprocess > /dev/null 2>&1
$FULL -k regex:"bilinear_crop|ssg_nms_merge|ssg_lowres" -c 3 -o gpurun_out/ev/ssgpost python tests/prof_kernels.py ssgpost > /dev/null 2>&1
$FULL -k regex:"peak_scan|jaccard|peak_select" -c 3 -o gpurun_out/ev/tail_blobs python tests/prof_kernels.py tail_blobs > /dev/null 2>&1
$FULL -k regex:"peak_scan" -c 1 -o gpurun_out/ev/tail_stress python tests/prof_kernels.py tail_stress > /dev/null 2>&1
# 4. the tensor-bound GEMM class and the attention kernel, refreshed
FULL2="ncu --clock-control none --set full --import-source on -f"
$FULL2 -k regex:gemm_tc -c 4 -o gpurun_out/ev/gemm_vis3 python tests/prof_gemm.py > /dev/null 2>&1
$FULL2 -k regex:attention_tc -c 1 -o gpurun_out/ev/attn_self python tests/prof_attn.py > /dev/null 2>&1
ls -la gpurun_out/ev | tail -20
