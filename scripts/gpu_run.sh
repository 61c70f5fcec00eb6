mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload tail --steps 10 > gpurun_out/bench_tail.json 2>gpurun_out/bench_tail.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tail.json')); print(d['blobs'], d['stress'])"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off --csv --log-file gpurun_out/tail_launches.csv python tests/prof_tail.py 4096 > gpurun_out/prof_tail.log 2>&1
