mkdir -p gpurun_out
python bench.py --workload tail --steps 30 > gpurun_out/tail.json 2> gpurun_out/tail.err; python -c "import json; d=json.loads(open('gpurun_out/tail.json').read().strip().splitlines()[-1]); print('tail graph', d['ms_per_step'], d['roofline']['frac'], d['stress']['ms'], d['blobs']['parity_spot_check'])"
python bench.py --workload tail --steps 30 --tail-no-graph > gpurun_out/tail2.json 2> gpurun_out/tail2.err; python -c "import json; d=json.loads(open('gpurun_out/tail2.json').read().strip().splitlines()[-1]); print('tail eager', d['ms_per_step'], d['roofline']['frac'], d['stress']['ms'], d['blobs']['parity_spot_check'])"
tail -2 gpurun_out/tail.err
