mkdir -p gpurun_out/lin
for c in 0 1 2 3 4 5 6 7 8 9 10; do python tests/prof_linear.py 43264 512 512 $c 2>&1 | tail -1; done
python tests/prof_linear.py 43264 1536 512 0 | tail -1
python tests/prof_linear.py 43264 1536 512 5 | tail -1
ncu --clock-control none --set full --import-source on -f --profile-from-start off -k regex:gemm_tc -c 1 -o gpurun_out/lin/q python tests/prof_linear.py 43264 512 512 0 > gpurun_out/lin/log.txt 2>&1
ncu -i gpurun_out/lin/q.ncu-rep --page source --csv --print-source sass > gpurun_out/lin/q_sass.csv 2>gpurun_out/lin/err.txt
ncu -i gpurun_out/lin/q.ncu-rep --page raw --csv > gpurun_out/lin/q_raw.csv 2>>gpurun_out/lin/err.txt
rm -f gpurun_out/lin/q.ncu-rep
