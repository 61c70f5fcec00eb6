mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest2.log; tail -3 gpurun_out/r2_pytest2.log
python bench.py --steps 30 --warmup 3 --dump-ops gpurun_out/r2_ops_b.txt > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 400 gpurun_out/r2_bench_b.err
for T in 0 8 16 24 32; do
  CROG_TEXT_SMS=$T python bench.py --steps 30 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_q_t$T.json 2>gpurun_out/r2_q_t$T.err
  python -c "import json;d=json.loads(open('gpurun_out/r2_q_t$T.json').read().strip().splitlines()[-1]);print('TEXT_SMS',$T,round(d['value'],1),round(d['ms_per_step'],3),d['clocks']['sm_mhz'])"
done
CROG_STEM_PAIRS=0 python bench.py --steps 30 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --dump-ops gpurun_out/r2_ops_nopairs.txt > gpurun_out/r2_q_nopairs.json 2>gpurun_out/r2_q_nopairs.err
python -c "import json;d=json.loads(open('gpurun_out/r2_q_nopairs.json').read().strip().splitlines()[-1]);print('NOPAIRS',round(d['value'],1),round(d['ms_per_step'],3))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_tail_launches_raw.csv python bench.py --workload tail --steps 2 --tail-no-graph > gpurun_out/r2_tail_ncu.json 2> gpurun_out/r2_tail_ncu.err
python bench.py --workload tail --steps 20 > gpurun_out/r2_tail_b.json 2>gpurun_out/r2_tail_b.err; python -c "import json;d=json.loads(open('gpurun_out/r2_tail_b.json').read().strip().splitlines()[-1]);print('TAIL blobs',d['blobs'],'stress',d['stress'])"
