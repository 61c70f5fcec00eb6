mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 3 --dump-ops gpurun_out/ops.txt > gpurun_out/bench_fwd.json 2>gpurun_out/bench_fwd.err; cat gpurun_out/bench_fwd.json | cut -c1-1500
