mkdir -p gpurun_out
for rep in 1 2; do
CROG_AUTOTUNE_MAX_CFG=9 python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b0.json 2> gpurun_out/b0.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b0.json').read().strip().splitlines()[-1]); print('no E12', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('E12   ', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done
