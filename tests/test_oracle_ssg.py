"""CPU tests of the SSG oracle (config 4): the restatement against the golden vectors written from the real reference
(oracle/make_golden_ssg.py), the Gaussian against the installed scipy, and the host side of the SSG drop-in."""
import os

import numpy as np
import pytest
import torch

from crog_b200 import synth
from oracle import ssg_forward as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_forward_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "ssg_perturbed_288.npz"))
    size, batch = int(g["size"]), int(g["batch"])
    cfg = synth.ssg_cfg(img_size=size)
    sd = synth.make_ssg_state_dict(cfg, 0, "perturbed")
    rgb, depth = synth.make_ssg_inputs(batch, size)
    out, inter = O.ssg_forward(sd, cfg, rgb, depth, keep=True)
    assert len(out["anchors"]) == out["cls_pred"].shape[1] * 4
    tol = 2e-5
    assert np.abs(out["protos"][:, ::4, ::4].numpy() - g["protos_s"]).max() <= tol
    assert np.abs(out["cls_pred"][:, ::37].numpy() - g["cls_s"]).max() <= tol
    assert np.abs(out["box_pred"][:, ::37].numpy() - g["box_s"]).max() <= tol
    assert np.abs(out["ins_coef_pred"][:, ::37].numpy() - g["coef_s"]).max() <= tol
    assert np.abs(out["grasp_coef_pred"][:, ::37].numpy() - g["gcoef_s"]).max() <= tol
    assert np.abs(inter["c5"][:, ::64].numpy() - g["c5_s"]).max() <= tol
    assert np.abs(inter["p7"][:, ::8].numpy() - g["p7_s"]).max() <= tol


def test_gaussian_is_scipy_bit_for_bit():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for shape in ((37, 53), (480, 640), (5, 9)):
        img = rng.random(shape, dtype=np.float32)
        img[rng.random(shape) < 0.3] = 0.0  # cropped regions are exact zeros
        want = ndi.gaussian_filter(img, 2.0, mode="nearest", truncate=4.0)
        got = O.gaussian_f32(img, 2.0)
        assert got.dtype == np.float32 and np.array_equal(got, want), shape


def test_post_processing_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "ssg_post_s6.npz"))
    cfg = synth.ssg_cfg()
    od = synth.make_ssg_output_dict(cfg, n_confident=8, seed=int(g["seed"]))
    out = O.ssg_post_processing(cfg, od, {"ori_size": (480, 640)}, keep=True)
    assert np.array_equal(out["cls"], g["cls"])
    assert np.array_equal(out["bboxes"], g["bboxes"])
    assert np.array_equal(out["ins_masks"].sum((1, 2)), g["ins_area"])
    assert np.array_equal(out["grasp_masks"][0].sum(-1).astype(np.float32), g["qua_rowsum"])
    g5 = g["grasps_top5"]
    for i, rows in enumerate(out["grasps_top5"]):
        assert len(rows) == int(np.isfinite(g5[i, :, 0]).sum())
        for j, r in enumerate(rows):
            assert np.array_equal(np.asarray(r, np.float64), g5[i, j])
        assert out["grasps_top1"][i] == rows[:1]


def test_fast_nms_semantics():
    """Degenerate boxes give IoU 0/0 = NaN, which (like torch.max) suppresses the later box; ties keep index order."""
    cfg = synth.ssg_cfg(top_k=4, max_detections=3)
    box = torch.tensor([[0.1, 0.1, 0.5, 0.5], [0.1, 0.1, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9], [0.2, 0.2, 0.2, 0.2], [0.7, 0.1, 0.9, 0.3]])
    cls = torch.tensor([[0.9, 0.9, 0.8, 0.7, 0.6], [0.1, 0.2, 0.3, 0.95, 0.05]])
    cid, sc, bx, _, _ = O.fast_nms(cfg, box, cls, torch.zeros(5, 32), torch.zeros(5, 4, 32))
    # class 0 keeps top-4 {0,1,2,3}: 1 duplicates 0 -> dropped, 3 is degenerate but only later columns see it; class 1 top-4 = {3,2,1,0}:
    # column of anchor 2 sees the NaN from degenerate anchor 3 -> dropped, as are 1 and 0
    assert sc.tolist() == pytest.approx([0.95, 0.9, 0.8])
    assert cid.tolist() == [1, 0, 0]


def test_ssg_module_host_side():
    from crog_b200 import _lib as L
    from crog_b200.model import SSG, build_ssg

    cfg = synth.ssg_cfg()
    model, params = build_ssg(cfg)
    assert isinstance(model, SSG) and len(list(params)) > 0
    sd = synth.make_ssg_state_dict(cfg, 0, "init")
    model.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)  # DataParallel-style checkpoint keys
    assert len(model.state_dict()) == len(sd) == 356
    assert len(model.anchors) == 18525 * 4
    rgb, depth = synth.make_ssg_inputs(1, 544)
    with pytest.raises(L.CrogError):
        model({"rgb": rgb, "depth": depth})  # no CPU fallback
