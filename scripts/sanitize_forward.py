"""One eager bf16 forward (batch 1, no CUDA graph, no autotune) + decode, for `compute-sanitizer --tool memcheck`."""
import os, sys
os.environ["CROG_AUTOTUNE"] = "0"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.engine import GraspEvaluator
from crog_b200.model import CROG
cfg = synth.default_cfg(17)
model = CROG(cfg, precision="bf16", use_cuda_graph=False)
model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed"))
model = model.cuda()
model.autotune = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
img, word = synth.make_inputs(B, 17)
gt, cnt = synth.make_gt_rects(B, 64, seed=4)
ev = GraspEvaluator(model)
out = ev.step(img.cuda(), word.cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(cnt).cuda())
torch.cuda.synchronize()
print("sanitize_forward: done", int(out[2][0]))
