"""Randomised sweep of crog_sigmoid_bicubic against torch (sigmoid + F.interpolate bicubic, align_corners=True): random input /
output extents (up- and down-sampling, so both the tiled kernel and its generic fallback run), batch sizes and plane counts.
python scripts/fuzz_glue.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.engine import postprocess

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad, t0 = 0, time.time()
for t in range(trials):
    B = int(rng.integers(1, 5)); NP = int(rng.choice([1, 5]))
    h, w = int(rng.integers(2, 130)), int(rng.integers(2, 130))
    f = rng.choice([0.5, 1.0, 2.0, 3.3, 4.0, 4.0, 7.7])
    H, W = max(2, int(h * f) + int(rng.integers(0, 3))), max(2, int(w * f) + int(rng.integers(0, 3)))
    torch.manual_seed(int(rng.integers(1 << 30)))
    maps = [torch.randn(B, 1, h, w, device="cuda") * 3 for _ in range(NP)]
    got = postprocess(maps, (H, W))
    torch.cuda.synchronize()
    err = 0.0
    for i, m in enumerate(maps):
        sig = (i in (0, 1, 4)) if NP == 5 else True
        x = torch.sigmoid(m) if sig else m
        want = F.interpolate(x, size=(H, W), mode="bicubic", align_corners=True)[:, 0]
        err = max(err, float((got[i] - want).abs().max()) / max(1.0, float(want.abs().max())))
    if not err < 3e-5:
        bad += 1; print("MISMATCH", dict(B=B, NP=NP, h=h, w=w, H=H, W=W), err)
print(f"{trials} trials, {bad} mismatching, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
