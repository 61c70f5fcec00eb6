mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:peak_scan_kernel -s 10 -c 1 -f -o gpurun_out/r2_scan_stress python bench.py --workload tail --steps 2 --tail-no-graph > gpurun_out/r2_scan_ncu.json 2> gpurun_out/r2_scan_ncu.err
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_tail_launches_raw3.csv -k regex:peak_ python bench.py --workload tail --steps 2 --tail-no-graph > /dev/null 2>&1
