"""Randomised sweep of the two OpenCV-exact warps against the numpy oracle (pinned to cv2 by tests/golden/warp_cases.npz):
random affine matrices (rotation, anisotropic scale, shear, translations that push the footprint over every border),
random source / destination sizes, random border values; float32 maps and uint8 frames (letterbox + normalisation).
Bit-exact.  python scripts/fuzz_warp.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.utils import warp as W
from oracle import warp_affine as WA

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad, t0 = 0, time.time()
for t in range(trials):
    Hs, Ws = int(rng.integers(4, 90)), int(rng.integers(4, 90))
    h, w = int(rng.integers(1, 100)), int(rng.integers(1, 100))
    th = rng.uniform(-np.pi, np.pi) if rng.random() < 0.6 else 0.0
    sx, sy = rng.uniform(0.3, 3.0), rng.uniform(0.3, 3.0)
    sh = rng.uniform(-0.3, 0.3) if rng.random() < 0.3 else 0.0
    A = np.array([[np.cos(th) * sx, -np.sin(th) * sy + sh], [np.sin(th) * sx, np.cos(th) * sy]])
    M = np.concatenate([A, rng.uniform(-30, 60, (2, 1))], 1).astype(np.float64)
    bv = float(rng.choice([0.0, 0.0, 0.25, -1.5]))
    if rng.random() < 0.6:  # float32 maps, NP planes x B samples with one matrix
        NP, B = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        src = rng.standard_normal((NP, B, Hs, Ws)).astype(np.float32)
        got = W.warp_affine_cubic(torch.from_numpy(src).cuda(), M, (w, h), bv).cpu().numpy()
        ok = all(np.array_equal(got[p, b], WA.warp_affine_cubic_f32(src[p, b], M, (w, h), bv)) for p in range(NP) for b in range(B))
        what = "f32"
    else:           # uint8 frames -> normalised network input
        B = int(rng.integers(1, 4))
        img = rng.integers(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
        got = W.preprocess_images(torch.from_numpy(img).cuda(), M, (h, w)).cpu().numpy()
        ok = all(np.array_equal(got[b], WA.preprocess_image(img[b], M, (h, w))) for b in range(B))
        what = "u8"
    if not ok:
        bad += 1; print("MISMATCH", what, dict(Hs=Hs, Ws=Ws, h=h, w=w), M.tolist(), bv)
print(f"{trials} trials, {bad} mismatching, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
