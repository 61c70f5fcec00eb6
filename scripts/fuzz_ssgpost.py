"""Randomised sweep of the batched SSG post-processing against the CPU oracle (the reference's ssg_post_processing
restated and pinned, oracle/ssg_forward.py): random numbers of confident instances, seeds and original sizes; detections
(classes, order) exact, boxes 1e-3 px, masks 1e-5, decoded peak positions equal.  (At least one confident instance per
image: the reference - and hence the oracle - raises in fast_nms when no anchor passes the score threshold; the drop-in
returns zero detections there.)  python scripts/fuzz_ssgpost.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.utils import grasp_eval as GE
from oracle import ssg_forward as O

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cfg = synth.ssg_cfg()
bad, t0, nimg = 0, time.time(), 0
for t in range(trials):
    B = int(rng.integers(1, 5))
    ori = (480, 640) if rng.random() < 0.5 else (int(rng.integers(60, 500)), int(rng.integers(60, 640)))
    ods = [synth.make_ssg_output_dict(cfg, n_confident=int(rng.integers(1, 13)), seed=int(rng.integers(1 << 20))) for _ in range(B)]
    batch = {k: torch.cat([od[k] for od in ods]).cuda() for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
    batch["anchors"] = ods[0]["anchors"]
    got = GE.ssg_post_processing_batched(cfg, batch, ori)
    torch.cuda.synchronize()
    for od, g in zip(ods, got):
        nimg += 1
        ref = O.ssg_post_processing(cfg, od, {"ori_size": ori}, keep=True)
        n = len(ref["cls"])
        ok = g["n"] == n and np.array_equal(g["cls"].cpu().numpy(), ref["cls"])
        if ok and n:
            hr = g["hr"].cpu().numpy()
            ok = ok and np.abs(g["boxes"].cpu().numpy() * ori[1] - ref["bboxes"]).max() <= 1e-3  # the reference scales all four by ori_w
            ok = ok and np.abs(hr[0] - ref["ins_masks"]).mean() <= 1e-5
            ok = ok and np.abs(hr[1] - ref["grasp_masks"][0]).max() <= 1e-5 and np.abs(hr[4] - ref["grasp_masks"][2]).max() <= 1e-5
            npk, gr = g["n_peaks"].cpu().numpy(), g["grasps"].cpu().numpy()
            for i in range(n):
                a = [(r[0], r[1]) for r in gr[i, :npk[i]].tolist()]
                b = [(r[0], r[1]) for r in ref["grasps_top5"][i]]
                ok = ok and a == b
        if not ok:
            bad += 1; print("MISMATCH trial", t, "ori", ori, "n", g["n"], n)
print(f"{trials} trials, {nimg} images, {bad} mismatching, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
