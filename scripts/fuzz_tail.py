"""Randomised parity sweep of crog_detect_grasps (+ crog_jaccard) against the C oracle: map sizes (widths that are and are not
multiples of 4, maps smaller than the 5x5 window), K, thresholds and value distributions (iid, quantised to few levels,
blobs, constant, sparse spikes, row / column ramps with ties).  Not part of the pytest suite (minutes of oracle time);
run once per kernel change:  python scripts/fuzz_tail.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.utils import grasp_eval as GE
from oracle import grasp_tail_c as TC

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda", 0)


def make(kind, n, H, W):
    if kind == "iid":
        return rng.random((n, H, W), dtype=np.float32)
    if kind == "quant":
        lv = int(rng.integers(2, 65))
        return (np.floor(rng.random((n, H, W), dtype=np.float32) * lv) / lv).astype(np.float32)
    if kind == "const":
        return np.full((n, H, W), float(rng.choice([0.0, 0.5, 1.0])), np.float32)
    if kind == "spikes":
        q = np.zeros((n, H, W), np.float32)
        m = rng.random((n, H, W)) < 0.01
        q[m] = rng.random(int(m.sum()), dtype=np.float32) * 0.6 + 0.4
        return q
    if kind == "ramp":
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        base = (xx / max(W - 1, 1)) if rng.random() < 0.5 else (yy / max(H - 1, 1))
        return np.repeat(base[None], n, 0).astype(np.float32)
    if kind == "blobs":
        S = max(H, W, 32)
        q = synth.make_tail_maps(n, "blobs", seed=int(rng.integers(1 << 30)), size=S)[0]
        return np.ascontiguousarray(q[:, :H, :W])
    raise ValueError(kind)


bad = 0
t0 = time.time()
for t in range(trials):
    H = int(rng.choice([3, 5, 6, 9, 17, 33, 64, 104, 131, 200, 416]))
    W = int(rng.choice([4, 5, 8, 12, 31, 64, 100, 120, 124, 128, 241, 416, 480, 500]))
    K = int(rng.choice([1, 2, 5, 9, 32]))
    thr = float(rng.choice([0.0, 0.4, 0.4, 0.9]))
    kind = str(rng.choice(["iid", "quant", "const", "spikes", "ramp", "blobs", "iid", "quant"]))
    n = 6 if H * W > 50000 else 16
    q = make(kind, n, H, W)
    s = rng.standard_normal((n, H, W)).astype(np.float32); c = rng.standard_normal((n, H, W)).astype(np.float32)
    w = rng.random((n, H, W), dtype=np.float32)
    peaks, cnt, grasps = GE.detect_grasps_batched(*[torch.from_numpy(a).to(dev) for a in (q, s, c, w)], K, threshold=thr)
    torch.cuda.synchronize()
    peaks, cnt, grasps = peaks.cpu().numpy(), cnt.cpu().numpy(), grasps.cpu().numpy()
    for b in range(n):
        ref = TC.peak_local_max(q[b], thr, K)
        ok = cnt[b] == len(ref) and np.array_equal(peaks[b, :cnt[b]], ref.astype(np.int32)) and (peaks[b, cnt[b]:] == -1).all()
        if not ok:
            bad += 1
            print("MISMATCH trial", t, dict(H=H, W=W, K=K, thr=thr, kind=kind, map=b, got=peaks[b, :cnt[b]].tolist(), want=ref.tolist()))
            break
print(f"{trials} trials, {bad} mismatching configurations, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
