mkdir -p gpurun_out
python tests/prof_attn.py
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -2
