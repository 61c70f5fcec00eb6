mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops.txt 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'])"
grep -E " stem.conv2 | stem.conv3 |layer1.1.conv2 | proj.vis.3 | proj.vis.1 |f2_v_proj" gpurun_out/ops.txt | awk '{printf "   %s %s\n", $2, $3}'
