mkdir -p gpurun_out/st
ncu --clock-control none --set full --import-source on -f -k regex:peak_scan -s 3 -c 1 -o gpurun_out/st/iid python scripts/time_tail.py iid > gpurun_out/st/log.txt 2>&1
ncu -i gpurun_out/st/iid.ncu-rep --page source --csv --print-source sass > gpurun_out/st/iid_sass.csv 2>gpurun_out/st/err.txt
ncu -i gpurun_out/st/iid.ncu-rep --page raw --csv > gpurun_out/st/iid_raw.csv 2>>gpurun_out/st/err.txt
rm -f gpurun_out/st/iid.ncu-rep
