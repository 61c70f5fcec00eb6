"""Small invocations of the memory-bound / tail / SSG kernels at ragged sizes, meant to run under
`compute-sanitizer --tool memcheck` (no tcgen05 GEMMs here: the sanitizer is very slow on them):
python scripts/sanitize_small.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L
from crog_b200 import synth
from crog_b200.engine import postprocess
from crog_b200.utils import grasp_eval as GE
from crog_b200.utils import warp as W

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
# decode: aligned fast path, odd width, unaligned base, tiny maps
for H, Wd, off in ((104, 128, 0), (37, 53, 0), (48, 64, 1), (5, 8, 0), (64, 416, 0)):
    n = 3
    buf = torch.rand(n * H * Wd + 8, device=dev)
    q = buf[off:off + n * H * Wd].view(n, H, Wd)
    s, c, w = torch.randn(n, H, Wd, device=dev), torch.randn(n, H, Wd, device=dev), torch.rand(n, H, Wd, device=dev)
    peaks, cnt, grasps = GE.detect_grasps_batched(q, s, c, w, 5)
    gt, gcnt = synth.make_gt_rects(n, 64, seed=4)
    GE.jacquard_batched(grasps, cnt, torch.from_numpy(gt).to(dev), torch.from_numpy(gcnt).to(dev))
# Gaussian: one-kernel and two-pass forms at ragged sizes
for H, Wd in ((97, 133), (33, 7), (480, 640)):
    m = torch.rand(2, H, Wd, device=dev)
    GE.gaussian_batched(m, 2.0)
    GE.gaussian_batched(m, 2.0, out=m)
# glue: tiled and generic
for h, w_, Hh, Ww in ((104, 104, 416, 416), (13, 29, 50, 97), (40, 40, 20, 20)):
    postprocess([torch.randn(2, 1, h, w_, device=dev) for _ in range(5)], (Hh, Ww))
# warps
M = np.array([[0.8, 0.1, 3.0], [-0.1, 0.9, 2.0]])
W.warp_affine_cubic(torch.rand(2, 3, 37, 41, device=dev), M, (53, 29), 0.0)
W.preprocess_images(torch.randint(0, 256, (2, 37, 41, 3), dtype=torch.uint8, device=dev), M, (29, 53))
mat, mat_inv = W.get_transform_mat((480, 640), (416, 416), inverse=True)
W.preprocess_images(torch.randint(0, 256, (1, 480, 640, 3), dtype=torch.uint8, device=dev), mat, (416, 416))
W.warp_affine_cubic(torch.rand(5, 1, 416, 416, device=dev), mat_inv, (640, 480), 0.0)
# SSG: stem patches at ragged sizes, batched post-processing at an odd original size
for S, cin, Kp in ((36, 4, 256), (150, 3, 192), (274, 4, 200)):
    O = (S - 1) // 2 + 1
    rgb, depth = torch.rand(2, 3, S, S, device=dev), torch.rand(2, 1, S, S, device=dev)
    out = torch.zeros((2 * O * O, Kp), device=dev, dtype=torch.bfloat16)
    L.check(L.lib().crog_stem7_patches(rgb.data_ptr(), depth.data_ptr() if cin == 4 else None, 2, S, S, cin, Kp, out.data_ptr(), L.BF16, L.stream_ptr()))
cfg = synth.ssg_cfg()
ods = [synth.make_ssg_output_dict(cfg, n_confident=n, seed=s_) for n, s_ in ((3, 6), (0, 7), (5, 8))]
batch = {k: torch.cat([od[k] for od in ods]).to(dev) for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
batch["anchors"] = ods[0]["anchors"]
GE.ssg_post_processing_batched(cfg, batch, (241, 323))
GE.ssg_post_processing_batched(cfg, batch, (480, 640))
torch.cuda.synchronize()
print("sanitize_small: done")
