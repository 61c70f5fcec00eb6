mkdir -p gpurun_out
python scripts/bf16_err.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 30 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('fused  ', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
grep -E "attnpool" gpurun_out/ops_tuned.txt
