mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 30 --warmup 3 --dump-ops gpurun_out/r2_ops_d.txt > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 400 gpurun_out/r2_bench_d.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),d['ms_per_step'],'e2e',round(d['e2e']['value'],1),'orig',round(d['e2e_variants']['original_resolution']['value'],1),'f32',round(d['e2e_variants']['fp32_tensors']['value'],1))
print('roof',d['roofline']['achieved'],d['roofline']['frac'],'step frac',d['roofline']['whole_step_frac_of_peak'])
print('tail blobs',d['tail']['blobs']); print('tail stress',d['tail']['stress'])
print('ssg',d['ssg']['samples_per_s'],d['ssg']['forward_ms'],d['ssg']['post_ms'],d['ssg']['roofline']['frac'])
print('parity',d['parity']['j_parity'],d['parity']['forward_ok'],d['parity']['j_counters_allreduced'])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_tail_launches_raw2.csv python bench.py --workload tail --steps 2 --tail-no-graph > gpurun_out/r2_tail_ncu.json 2> gpurun_out/r2_tail_ncu.err
grep -c peak_scan gpurun_out/r2_tail_launches_raw2.csv
