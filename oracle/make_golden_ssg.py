"""ORACLE tooling: pin oracle/ssg_forward.py against the real reference and write tests/golden/ssg_*.npz.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_ssg.py

1. Forward: imports the *unmodified* ``model.ssg.SSG`` from /root/reference, loads our synthetic state-dict
   with ``strict=True`` (which pins the name/shape table of crog_b200/spec.py:ssg_tensor_specs), runs it in
   eval mode on CPU fp32 and asserts oracle.ssg_forward agrees to fp32 round-off.
2. Post-processing: imports the *unmodified* ``utils/grasp_eval.py`` of the reference.  Its third-party
   imports that are absent offline (scikit-image, matplotlib) are satisfied with stub modules whose three
   functions are the oracle's restatements (gaussian, peak_local_max, polygon — SURVEY.md App. A), and the
   NumPy-1.x aliases it uses (np.int0 / np.float) are restored.  So every line of the reference's own
   ``ssg_post_processing`` / ``fast_nms`` / ``crop`` / ``box_iou`` / ``detect_grasps`` executes, and the
   oracle's restatement of them is asserted equal on a synthetic ``output_dict``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from crog_b200 import synth  # noqa: E402
from oracle import grasp_tail as T  # noqa: E402
from oracle import ssg_forward as O  # noqa: E402


def import_reference_grasp_eval():
    sk = types.ModuleType("skimage")
    sk_draw, sk_filters, sk_feature = types.ModuleType("skimage.draw"), types.ModuleType("skimage.filters"), types.ModuleType("skimage.feature")
    sk_draw.polygon = T.polygon
    sk_filters.gaussian = lambda img, sigma, preserve_range=True: O.gaussian_f32(img, sigma)
    sk_feature.peak_local_max = lambda img, min_distance=2, threshold_abs=0.4, num_peaks=5: T.peak_local_max(img, min_distance, threshold_abs, num_peaks)
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    for name, m in (("skimage", sk), ("skimage.draw", sk_draw), ("skimage.filters", sk_filters), ("skimage.feature", sk_feature),
                    ("matplotlib", mpl), ("matplotlib.pyplot", plt)):
        sys.modules.setdefault(name, m)
    if not hasattr(np, "int0"):
        np.int0 = np.int64
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    try:
        import utils.grasp_eval as ref_ge  # reference, unmodified
    finally:
        sys.path.remove(REF)
    return ref_ge


def forward_case(tag: str, mode: str, batch: int, size: int):
    cfg = synth.ssg_cfg(img_size=size)
    sys.path.insert(0, REF)
    try:
        from model.ssg import SSG as RefSSG  # reference
    finally:
        sys.path.remove(REF)
    ref = RefSSG(cfg).eval()
    sd = synth.make_ssg_state_dict(cfg, 0, mode)
    rsd = ref.state_dict()
    assert set(rsd) == set(sd), set(rsd) ^ set(sd)
    for k in sd:
        assert tuple(rsd[k].shape) == tuple(sd[k].shape), k
    ref.load_state_dict(sd, strict=True)
    rgb, depth = synth.make_ssg_inputs(batch, size)
    with torch.no_grad():
        r = ref({"rgb": rgb, "depth": depth})
    o, inter = O.ssg_forward(sd, cfg, rgb, depth, keep=True)
    assert r["anchors"] == o["anchors"]
    errs = {k: float((r[k] - o[k]).abs().max()) for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
    print(tag, {k: "%.1e" % v for k, v in errs.items()})
    assert all(v <= 2e-5 for v in errs.values()), errs
    out = {"size": np.int64(size), "batch": np.int64(batch), "mode": np.array(mode),
           "protos_s": r["protos"][:, ::4, ::4].numpy(), "cls_s": r["cls_pred"][:, ::37].numpy(), "box_s": r["box_pred"][:, ::37].numpy(),
           "coef_s": r["ins_coef_pred"][:, ::37].numpy(), "gcoef_s": r["grasp_coef_pred"][:, ::37].numpy(),
           "c3_s": inter["c3"][:, ::32, ::4, ::4].numpy(), "c5_s": inter["c5"][:, ::64].numpy(), "p3_s": inter["p3"][:, ::16, ::4, ::4].numpy(),
           "p7_s": inter["p7"][:, ::8].numpy()}
    path = os.path.join(ROOT, "tests", "golden", f"ssg_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def post_case(tag: str, seed: int):
    ref_ge = import_reference_grasp_eval()
    cfg = synth.ssg_cfg()
    od = synth.make_ssg_output_dict(cfg, n_confident=8, seed=seed)
    dd = {"ori_size": (480, 640)}
    r = ref_ge.ssg_post_processing(cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in od.items()}, dd)
    o = O.ssg_post_processing(cfg, od, dd, keep=True)
    assert np.array_equal(r["cls"], o["cls"]), (r["cls"], o["cls"])
    assert np.array_equal(r["bboxes"], o["bboxes"])
    assert np.array_equal(r["ins_masks"], o["ins_masks"])
    (rq, ra, rw), (oq, oa, ow) = r["grasp_masks"], o["grasp_masks"]
    assert np.array_equal(np.asarray(rq), oq) and np.array_equal(np.asarray(rw), ow)
    # the reference's angle map is libm atan2f (not bit-reproducible across CPUs); the oracle defines it as
    # float32(atan2(double, double)) * 0.5f.  NumPy's SIMD atan2f is accurate to a few ulp,
    # so agreement within 4 float32 ulp is what is asserted here.
    ra = np.asarray(ra, np.float32)
    assert np.all(np.abs(ra - oa) <= 4 * np.spacing(np.maximum(np.abs(ra), np.abs(oa)).astype(np.float32)))
    assert len(r["grasps_top5"]) == len(o["grasps_top5"])
    # Under NumPy 2 the unmodified reference computes width*100 and angle/pi*180 in float32 (no value-based
    # promotion); the oracle follows the pinned NumPy 1.24.3 (float64).  Peak positions must
    # agree exactly, the two float columns to float32 round-off.
    n_g = 0
    for ga, gb in zip(r["grasps_top5"], o["grasps_top5"]):
        assert len(ga) == len(gb)
        for x, y in zip(ga, gb):
            assert x[0] == y[0] and x[1] == y[1] and x[3] == y[3]
            assert abs(float(x[2]) - y[2]) <= 1e-4 * max(1.0, abs(y[2])) and abs(float(x[4]) - y[4]) <= 1e-4 * max(1.0, abs(y[4]))
            n_g += 1
    for ga, gb in zip(r["grasps_top1"], o["grasps_top1"]):
        assert len(ga) == len(gb) and all(x[0] == y[0] and x[1] == y[1] for x, y in zip(ga, gb))
    print(tag, "instances", len(o["cls"]), "grasps", n_g, "classes", o["cls"].tolist())
    assert len(o["cls"]) >= 4 and n_g >= 4, "synthetic output_dict should yield confident instances with grasps"
    g5 = np.full((len(o["cls"]), 5, 5), np.nan)
    for i, g in enumerate(o["grasps_top5"]):
        for j, row in enumerate(g):
            g5[i, j] = row
    out = {"seed": np.int64(seed), "cls": o["cls"], "bboxes": o["bboxes"], "scores": o["_debug"]["scores"], "grasps_top5": g5,
           "ins_area": o["ins_masks"].sum((1, 2)), "qua_rowsum": o["grasp_masks"][0].sum(-1).astype(np.float32),
           "wid_rowsum": o["grasp_masks"][2].sum(-1).astype(np.float32)}
    path = os.path.join(ROOT, "tests", "golden", f"ssg_post_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    post_case("s6", 6)
    forward_case("perturbed_288", "perturbed", 2, 288)
    forward_case("init_544", "init", 1, 544)
