q() { python bench.py --steps 40 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_q_$1.json 2>gpurun_out/r2_q_$1.err; python -c "import json;d=json.loads(open('gpurun_out/r2_q_$1.json').read().strip().splitlines()[-1]);print('$1',round(d['value'],1),'step',round(d['ms_per_step'],3),'fwd',round(d['forward_ms_per_step'],3),d['clocks']['sm_mhz'])"; }
q base1
CROG_B200_SO=$PWD/crog_b200/lib/libcrog_b200.hint.so q hint1
q base2
CROG_B200_SO=$PWD/crog_b200/lib/libcrog_b200.hint.so q hint2
