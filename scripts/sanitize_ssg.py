"""One eager bf16 SSG forward (288 x 288, batch 1, no CUDA graph, no autotune) for `compute-sanitizer --tool memcheck`."""
import os, sys
os.environ["CROG_AUTOTUNE"] = "0"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.model import SSG
cfg = synth.ssg_cfg(img_size=288)
model = SSG(cfg, precision="bf16", use_cuda_graph=False)
model.load_state_dict(synth.make_ssg_state_dict(cfg, 0, "perturbed"), strict=True)
model = model.cuda()
rgb, depth = synth.make_ssg_inputs(1, 288)
out = model({"rgb": rgb.cuda(), "depth": depth.cuda()})
torch.cuda.synchronize()
print("sanitize_ssg: done", tuple(out["protos"].shape))
