mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q -k "engine or second_operand" 2>&1 | tail -3
python bench.py --steps 50 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
