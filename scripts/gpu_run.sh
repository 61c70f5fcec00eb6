mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
for rep in 1 2; do
CROG_TEXT_PDL=0 python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b0.json 2> gpurun_out/b0.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b0.json').read().strip().splitlines()[-1]); print('no pdl', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('pdl   ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
