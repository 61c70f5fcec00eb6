mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "folded or tile_cfgs or plain_gemm or epilogue" 2>&1 | tail -5
CROG_FFN_LN_FOLD=0 python scripts/bf16_err.py 2>&1 | tail -1
CROG_FFN_LN_FOLD=1 python scripts/bf16_err.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CROG_FFN_LN_FOLD=0 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_nofold.json 2> gpurun_out/bench_nofold.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_nofold.json').read().strip().splitlines()[-1]); print('nofold', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --dump-ops gpurun_out/ops_tuned.txt > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_tuned.json').read().strip().splitlines()[-1]); print('fold', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'])"
grep -E "decoder.layers.0.ffn" gpurun_out/ops_tuned.txt
