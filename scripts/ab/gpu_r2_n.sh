python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 40 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --dump-ops gpurun_out/r2_ops_n.txt > gpurun_out/r2_q_n.json 2>gpurun_out/r2_q_n.err; python -c "import json;d=json.loads(open('gpurun_out/r2_q_n.json').read().strip().splitlines()[-1]);print('value',round(d['value'],1),'step',round(d['ms_per_step'],3),'fwd',round(d['forward_ms_per_step'],3),'roof',round(d['roofline']['frac'],4))"
grep "proj.gather\|proj.dynconv\|stem" gpurun_out/r2_ops_n.txt | cut -c1-120
