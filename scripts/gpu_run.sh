mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in 0 1; do python tests/prof_gemm_shape.py 692224 256 64 $cfg; python tests/prof_gemm_shape.py 43264 1024 256 $cfg; done
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_tuned.json').read().strip().splitlines()[-1]); print('tuned', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline'].get('autotuned_layers'), d['clocks'])"
