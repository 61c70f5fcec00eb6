# Round-2 evidence, end-of-round kernels: forward launch list + full captures of the kernels changed in the last stretch.
# The .ncu-rep files are reduced to the judged metrics on the box (profiles/extract.py) and deleted.
mkdir -p gpurun_out/ev3
NCU="ncu --clock-control none --profile-from-start off"
FULL="$NCU --set full --import-source on -f"
CROG_NO_FORK=1 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --csv --log-file gpurun_out/ev3/fwd_raw.csv python tests/prof_forward.py 64 gpurun_out/ev3/fwd_ops.tsv > gpurun_out/ev3/fwd.log 2>&1
python profiles/launch_list.py gpurun_out/ev3/fwd_raw.csv gpurun_out/ev3/fwd_ops.tsv gpurun_out/ev3/r2_launches.csv
cap() {  # name, kernel regex, count, command...
  n=$1; k=$2; c=$3; shift 3
  $FULL -k regex:"$k" -c $c -o gpurun_out/ev3/$n "$@" > /dev/null 2>&1
  python profiles/extract.py gpurun_out/ev3/$n.ncu-rep gpurun_out/ev3/r2_ncu_full_$n.csv
  rm -f gpurun_out/ev3/$n.ncu-rep
}
cap tail_blobs "peak_scan|jaccard|peak_select" 3 python tests/prof_kernels.py tail_blobs
cap tail_stress "peak_scan" 1 python tests/prof_kernels.py tail_stress
cap glue "sigmoid_bicubic" 1 python tests/prof_kernels.py glue
cap gaussian "gaussian" 1 python tests/prof_kernels.py gaussian
cap ssgpost "bilinear_crop|ssg_nms_merge|ssg_nms_class|ssg_lowres" 4 python tests/prof_kernels.py ssgpost
ncu --clock-control none --set full -f -k regex:stem7_patches -c 1 -o gpurun_out/ev3/stem7 python scripts/ssg_ops.py > /dev/null 2>&1
python profiles/extract.py gpurun_out/ev3/stem7.ncu-rep gpurun_out/ev3/r2_ncu_full_stem7.csv; rm -f gpurun_out/ev3/stem7.ncu-rep
# tail launch list (both distributions), eager
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file gpurun_out/ev3/tail_blobs_raw.csv python tests/prof_kernels.py tail_blobs > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file gpurun_out/ev3/tail_stress_raw.csv python tests/prof_kernels.py tail_stress > /dev/null 2>&1
ls -la gpurun_out/ev3
