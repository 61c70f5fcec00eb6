python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tile_cfgs or conv3 or tap_mask" 2>&1 | tail -3
for c in 0 18; do python tests/prof_conv3.py 64 208 104 64 64 $c; python tests/prof_conv3.py 64 104 104 64 64 $c; python tests/prof_conv3.py 64 208 104 64 64 $c 0; done
