mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -5
python -m pytest tests/test_gpu_model.py tests/test_gpu_ssg.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --dump-ops gpurun_out/r2_ops_j.txt > gpurun_out/r2_q_band.json 2>gpurun_out/r2_q_band.err; tail -c 300 gpurun_out/r2_q_band.err
python -c "import json;d=json.loads(open('gpurun_out/r2_q_band.json').read().strip().splitlines()[-1]);print('BAND value',round(d['value'],1),'step',round(d['ms_per_step'],3),'fwd',round(d['forward_ms_per_step'],3),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],4))"
grep "tile_cfg 1[1-5]" gpurun_out/r2_ops_j.txt | cut -c1-150
