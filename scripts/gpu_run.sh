mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
CROG_SIDE_HELPERS=0 python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b0.json 2> gpurun_out/b0.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b0.json').read().strip().splitlines()[-1]); print('inline ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('helpers', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
tail -3 gpurun_out/b.err
