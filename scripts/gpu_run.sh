mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -k "stem or fp32 or bf16" 2>&1 | tail -3
python bench.py --steps 30 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'])"
grep -E "stem" gpurun_out/ops_tuned.txt
