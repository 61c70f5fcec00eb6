mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_tuned.json').read().strip().splitlines()[-1]); print('tuned', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline'].get('autotuned_layers'))"
