mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 --dump-ops gpurun_out/r2_ops_g.txt > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -c 300 gpurun_out/r2_bench_g.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),d['ms_per_step'],'fwd',d['forward_ms_per_step'],'e2e',round(d['e2e']['value'],1),'orig',round(d['e2e_variants']['original_resolution']['value'],1),'f32',round(d['e2e_variants']['fp32_tensors']['value'],1))
print('roof',d['roofline']['achieved'],d['roofline']['frac'],'step frac',d['roofline']['whole_step_frac_of_peak'])
print('tail blobs',d['tail']['blobs']); print('tail stress',d['tail']['stress'])
print('ssg',d['ssg']['samples_per_s'],d['ssg']['forward_ms'],d['ssg']['post_ms'],d['ssg']['roofline']['frac'])
print('parity',d['parity']['j_parity'],d['parity']['forward_ok'],d['parity']['j_counters_allreduced'])
P
for c in 2 4; do python bench.py --workload tail --steps 20 --tail-chunks $c > gpurun_out/r2_tail_c$c.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/r2_tail_c$c.json').read().strip().splitlines()[-1]);print('CHUNKS',$c,'blobs',round(d['blobs']['ms'],3),'stress',round(d['stress']['ms'],3))"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ssg_launches_raw.csv python bench.py --workload ssg --steps 1 --ssg-batch 16 > gpurun_out/r2_ssg_ncu.json 2> gpurun_out/r2_ssg_ncu.err
