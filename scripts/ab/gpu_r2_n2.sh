mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "dataparallel or ragged" > gpurun_out/r2_dp_test.log 2>&1
grep -v "CUDAEvent" gpurun_out/r2_dp_test.log | tail -40
