mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "jaccard or detect or pipelined or reference_signature" 2>&1 | tail -2
python bench.py --workload tail --steps 30 > gpurun_out/tail.json 2> gpurun_out/tail.err; python -c "import json; d=json.loads(open('gpurun_out/tail.json').read().strip().splitlines()[-1]); print('tail', d['ms_per_step'], d['roofline']['frac'], d['stress']['ms'], d['blobs']['parity_spot_check'])"
