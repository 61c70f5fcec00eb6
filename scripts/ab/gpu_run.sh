# One GPU-box validation pass (used with: gpurun --timeout 1500 -- 'bash scripts/gpu_run.sh'):
# the GPU parity suite, the smoke entry point, the default bench line and the tail micro-benchmark.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python -c "import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF/s', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
python bench.py --workload tail > gpurun_out/tail_final.json 2> gpurun_out/tail_final.err
python -c "import json; d=json.loads(open('gpurun_out/tail_final.json').read().strip().splitlines()[-1]); print('tail', d['ms_per_step'], d['roofline']['frac'], d['stress']['ms'], d['blobs']['parity_spot_check'])"
