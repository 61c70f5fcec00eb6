mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 30 --warmup 3 --dump-ops gpurun_out/r2_ops_k.txt > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; tail -c 300 gpurun_out/r2_bench_k.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_k.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),d['ms_per_step'],'fwd',d['forward_ms_per_step'],'e2e',round(d['e2e']['value'],1),'orig',round(d['e2e_variants']['original_resolution']['value'],1),'f32',round(d['e2e_variants']['fp32_tensors']['value'],1))
print('roof',d['roofline']['achieved'],d['roofline']['frac'],'step frac',d['roofline']['whole_step_frac_of_peak'], d['clocks'])
print('tail blobs',d['tail']['blobs']['ms'],d['tail']['blobs']['frac_of_hbm'],'stress',d['tail']['stress']['ms'],d['tail']['stress']['frac_of_hbm'])
print('ssg',d['ssg']['samples_per_s'],d['ssg']['forward_ms'],d['ssg']['post_ms'],d['ssg']['roofline']['frac'])
print('parity',d['parity']['j_parity'],d['parity']['forward_ok'],d['parity']['j_counters_allreduced'],d['parity']['forward_rel_l2_per_map'])
print('cpu',d['cpu_baseline']['value'])
P
