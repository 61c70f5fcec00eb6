mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['n_gpus'], d['j_counters'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
