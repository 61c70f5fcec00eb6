mkdir -p gpurun_out
q() { python bench.py --steps 30 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_q_$1.json 2>gpurun_out/r2_q_$1.err; python -c "import json;d=json.loads(open('gpurun_out/r2_q_$1.json').read().strip().splitlines()[-1]);print('$1',round(d['value'],1),'step',round(d['ms_per_step'],3),'fwd',round(d['forward_ms_per_step'],3),d['clocks']['sm_mhz'])"; }
q default
CROG_NO_FORK=1 q nofork
CROG_DEBUG_SKIP_TEXT=1 q skiptext
CROG_TEXT_SMS=0 q t0
CROG_TEXT_SMS=32 q t32
q default2
python -m pytest tests/test_gpu_ssg.py -m gpu -q 2>&1 | tail -2
python bench.py --workload ssg --steps 10 > gpurun_out/r2_ssg_b.json 2>gpurun_out/r2_ssg_b.err; python -c "import json;d=json.loads(open('gpurun_out/r2_ssg_b.json').read().strip().splitlines()[-1]);print('SSG',round(d['value'],1),d['forward_ms'],d['post_ms'])"
