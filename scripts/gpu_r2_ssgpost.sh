mkdir -p gpurun_out/ssgp
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/ssgp/raw.csv python tests/prof_kernels.py ssgpost > gpurun_out/ssgp/log.txt 2>&1
