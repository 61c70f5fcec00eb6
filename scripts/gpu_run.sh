mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
CROG_NO_FORK=1 ncu --set full --import-source on --clock-control none --profile-from-start off -c 14 -o gpurun_out/front_b8 -f python tests/prof_forward.py 8 > gpurun_out/prof_front.log 2>&1; tail -2 gpurun_out/prof_front.log
ls -la gpurun_out
