"""Randomised parity sweep of crog_jaccard against the C oracle: every (prediction, ground-truth) pixel count and the J@1 /
J@K flags, over rectangles that are tiny, thin, huge (slow path), far outside the canvas, axis-aligned and at the angle
gate's edges, with K in {1, 5, 9, 32}.  python scripts/fuzz_jaccard.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.utils import grasp_eval as GE
from oracle import grasp_tail_c as TC

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda", 0)


def rect(style):
    if style == 0:   # ordinary
        return [rng.uniform(0, 640), rng.uniform(0, 480), rng.uniform(0, 105), 20, rng.uniform(-90, 90)]
    if style == 1:   # tiny / degenerate
        return [rng.uniform(100, 300), rng.uniform(100, 300), rng.choice([0.0, 0.3, 1.0, 2.5]), 20, rng.uniform(-90, 90)]
    if style == 2:   # axis aligned and 45 degrees (ties in the scanline arithmetic)
        return [float(rng.integers(50, 400)), float(rng.integers(50, 400)), float(rng.integers(1, 100)), 20, float(rng.choice([-90, -45, 0, 45, 90]))]
    if style == 3:   # partly / far outside the canvas
        return [rng.uniform(-300, 900), rng.uniform(-300, 800), rng.uniform(0, 100), 20, rng.uniform(-90, 90)]
    return [rng.uniform(0, 640), rng.uniform(0, 480), rng.uniform(0, 100), 20, rng.uniform(-90, 90)]


bad, t0 = 0, time.time()
for t in range(trials):
    B, K, M = 12, int(rng.choice([1, 5, 5, 9, 32])), int(rng.choice([1, 7, 64]))
    gt = np.zeros((B, M, 6), np.float64)
    cnt = rng.integers(0, M + 1, B).astype(np.int32)
    cnt[0] = M
    grasps = np.zeros((B, K, 5), np.float64)
    n = rng.integers(0, K + 1, B).astype(np.int32)
    for b in range(B):
        for m in range(M):
            r = rect(int(rng.integers(0, 5)))
            gt[b, m] = [r[0], r[1], rng.uniform(-20, 140) if rng.random() < 0.2 else r[2], rng.uniform(5, 60), r[4], 0.0]
        for k in range(K):
            if rng.random() < 0.5 and cnt[b] > 0:  # near a ground-truth rectangle, angle around the gate's 30 degrees
                m = int(rng.integers(0, cnt[b]))
                grasps[b, k] = [gt[b, m, 0] + rng.uniform(-15, 15), gt[b, m, 1] + rng.uniform(-15, 15), rng.uniform(0, 105), 20,
                                gt[b, m, 4] + rng.choice([-30.5, -30, -29.5, 0, 29.5, 30, 30.5, rng.uniform(-40, 40)])]
            else:
                grasps[b, k] = rect(int(rng.integers(0, 5)))
    gt_dev = torch.from_numpy(gt.copy()).to(dev)
    flags, inter, uni = GE.jacquard_batched(torch.from_numpy(grasps).to(dev), torch.from_numpy(n).to(dev), gt_dev,
                                            torch.from_numpy(cnt).to(dev), want_counts=True)
    torch.cuda.synchronize()
    flags, inter, uni = flags.cpu().numpy(), inter.cpu().numpy(), uni.cpu().numpy()
    ok = True
    for b in range(B):
        g_ref = gt[b, :cnt[b]].copy()
        jk = TC.jacquard(grasps[b, :n[b]], g_ref) if n[b] and cnt[b] else 0
        j1 = TC.jacquard(grasps[b, :1], gt[b, :cnt[b]].copy()) if n[b] and cnt[b] else 0
        if (flags[b, 0], flags[b, 1]) != (j1, jk):
            ok = False; print("FLAG MISMATCH", t, b, flags[b].tolist(), (j1, jk)); break
        if n[b] and cnt[b]:
            for k in range(n[b]):
                for m in range(cnt[b]):
                    if (inter[b, k, m], uni[b, k, m]) != TC.iou_counts(grasps[b, k], g_ref[m]):
                        ok = False; print("COUNT MISMATCH", t, b, k, m, grasps[b, k].tolist(), g_ref[m].tolist(), (inter[b, k, m], uni[b, k, m]), TC.iou_counts(grasps[b, k], g_ref[m])); break
                if not ok: break
        if not ok: break
    bad += 0 if ok else 1
print(f"{trials} trials, {bad} mismatching, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
