mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -k "sigmoid or engine or module_contract" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sigmoid_bicubic --csv python -c "
import torch, sys
sys.path.insert(0,'.')
from crog_b200.engine import postprocess
maps=[torch.randn(64,1,104,104,device='cuda') for _ in range(5)]
for _ in range(3): postprocess(maps,(416,416))
torch.cuda.synchronize()
" 2>&1 | grep -E "sigmoid" | tail -3

