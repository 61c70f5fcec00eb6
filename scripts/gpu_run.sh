mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops.txt 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'])"
grep -E "stem.conv1|up1|up2|f5_up|fq5_up|layer2.*conv2" gpurun_out/ops.txt
