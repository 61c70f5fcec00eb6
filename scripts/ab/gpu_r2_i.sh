mkdir -p gpurun_out
python -m pytest tests/test_gpu_warp.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3
python - <<'P'
import torch, time, sys
sys.path.insert(0,'.')
from crog_b200 import synth
from crog_b200.utils import warp as WP
dev=torch.device('cuda',0)
B=64
frames=synth.make_frames_u8(B).to(dev)
mat,mat_inv=WP.get_transform_mat((480,640),(416,416),inverse=True)
a=WP.device_affine(mat,B,dev); ai=WP.device_affine(mat_inv,B,dev)
post=torch.rand((5,B,416,416),device=dev)
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
print('preprocess us', round(t(lambda: WP.preprocess_images(frames,a,(416,416))),1))
print('warp us', round(t(lambda: WP.warp_affine_cubic(post,ai,(640,480),0.0)),1))
P
python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_q_e2e.json 2>gpurun_out/r2_q_e2e.err; python -c "import json;d=json.loads(open('gpurun_out/r2_q_e2e.json').read().strip().splitlines()[-1]);print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'orig',round(d['e2e_variants']['original_resolution']['value'],1),'f32',round(d['e2e_variants']['fp32_tensors']['value'],1))"
