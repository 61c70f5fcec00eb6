mkdir -p gpurun_out
for rep in 1 2; do
for so in crog_b200/lib/libcrog_b200.so crog_b200/lib/libcrog_nb.so; do
echo "== $so"
CROG_B200_SO=$PWD/$so python tests/prof_gemm_shape.py 692224 256 64 0; CROG_B200_SO=$PWD/$so python tests/prof_gemm_shape.py 43264 1024 256 0
CROG_B200_SO=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "import sys,json; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done; done
