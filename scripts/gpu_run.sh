mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_tuned.json').read().strip().splitlines()[-1]); print('tuned', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline'].get('autotuned_layers'), d['clocks'])"
CROG_AUTOTUNE=0 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_heur.json 2> gpurun_out/bench_heur.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_heur.json').read().strip().splitlines()[-1]); print('heur', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
