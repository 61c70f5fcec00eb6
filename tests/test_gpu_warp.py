"""GPU parity of the OpenCV-exact letterbox warps (rows f-1 / f-2): crog_warp_affine_cubic_f32, crog_preprocess_u8 and
crog_mask_iou through the C ABI against the cv2-pinned oracle and the cv2-generated golden vectors — bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

from crog_b200 import synth
from crog_b200.utils import warp as W

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "warp_cases.npz"))


def test_warp_f32_golden_small(g):
    for i in range(int(g["n_f32"])):
        src = torch.from_numpy(g[f"f32_{i}_src"]).cuda()[None]
        h, w = g[f"f32_{i}_out"].shape
        got = W.warp_affine_cubic(src, g[f"f32_{i}_M"], (w, h), float(g[f"f32_{i}_bv"]))[0].cpu().numpy()
        assert np.array_equal(got, g[f"f32_{i}_out"]), (i, np.abs(got - g[f"f32_{i}_out"]).max())


def test_preprocess_u8_golden_small(g):
    for i in range(int(g["n_u8"])):
        img = torch.from_numpy(g[f"u8_{i}_img"]).cuda()[None]
        got = W.preprocess_images(img, g[f"u8_{i}_M"], tuple(int(v) for v in g[f"u8_{i}_size"]))[0].cpu().numpy()
        assert np.array_equal(got, g[f"u8_{i}_out"]), (i, np.abs(got - g[f"u8_{i}_out"]).max())


def test_letterbox_full_size_golden(g):
    rng = np.random.default_rng(int(g["lb_seed"]))
    maps = rng.random((5, 416, 416), dtype=np.float32)
    inv = W.warp_affine_cubic(torch.from_numpy(maps).cuda()[:, None], g["lb_mat_inv"], (640, 480), 0.0)[:, 0]
    assert _sha(inv.cpu().numpy()) == str(g["lb_inv_sha"])
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    pre = W.preprocess_images(torch.from_numpy(img).cuda()[None], g["lb_mat"], (416, 416))[0]
    assert _sha(pre.cpu().numpy()) == str(g["lb_pre_sha"])
    tgt416 = (rng.random((416, 416)) > 0.7).astype(np.float32)
    tgt = W.warp_affine_cubic(torch.from_numpy(tgt416).cuda()[None], g["lb_mat_inv"], (640, 480), 0.0)
    iou, counts = W.mask_iou(inv[0:1], tgt)
    assert float(iou[0]) == float(g["lb_iou"])


def test_warp_batched_per_sample_matrices_vs_oracle():
    """B samples x NP planes, a different random matrix per sample, odd sizes, border value != 0."""
    from oracle import warp_affine as WA

    rng = np.random.default_rng(3)
    NP, B, Hs, Ws, h, w = 3, 5, 57, 83, 61, 45
    src = rng.standard_normal((NP, B, Hs, Ws)).astype(np.float32)
    mats = []
    for _ in range(B):
        ang, s = rng.uniform(-3.14, 3.14), rng.uniform(0.4, 2.5)
        mats.append([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-40, 40)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-40, 40)]])
    mats = np.array(mats)
    got = W.warp_affine_cubic(torch.from_numpy(src).cuda(), mats, (w, h), 0.25).cpu().numpy()
    for b in range(B):
        for p in range(NP):
            ref = WA.warp_affine_cubic_f32(src[p, b], mats[b], (w, h), 0.25)
            assert np.array_equal(got[p, b], ref), (p, b)


def test_warp_degenerate_and_far_outside():
    from oracle import warp_affine as WA

    src = np.arange(12 * 9, dtype=np.float32).reshape(12, 9)
    for M in ([[1, 0, 500], [0, 1, 0]], [[0, 0, 0], [0, 0, 0]], [[1e-3, 0, 0], [0, 1e-3, 0]], [[1, 0, -0.5], [0, 1, 0.5]]):
        M = np.array(M, np.float64)
        got = W.warp_affine_cubic(torch.from_numpy(src).cuda()[None], M, (20, 16), 0.0)[0].cpu().numpy()
        assert np.array_equal(got, WA.warp_affine_cubic_f32(src, M, (20, 16), 0.0)), M


def test_preprocess_batch_vs_oracle():
    from oracle import warp_affine as WA

    rng = np.random.default_rng(9)
    B, Ho, Wo = 3, 120, 160
    img = rng.integers(0, 256, (B, Ho, Wo, 3), dtype=np.uint8)
    mat, _ = W.get_transform_mat((Ho, Wo), (104, 104), inverse=True)
    got = W.preprocess_images(torch.from_numpy(img).cuda(), mat, (104, 104)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], WA.preprocess_image(img[b], mat, (104, 104))), b


@pytest.mark.parametrize("ori", [(480, 640), (375, 501)])
def test_evaluator_original_resolution_matches_oracle(ori):
    """The real pipeline (crog_engine.py:446-527): model -> sigmoid/bicubic -> inverse letterbox to the original size ->
    mask IoU, detect_grasps, Jaccard.  Given the GPU maps at 416^2, everything downstream is bit-exact vs the oracle.
    (375 x 501: odd sizes, so the warped planes are not 16-byte aligned and the decode takes the generic scan.)"""
    from crog_b200.engine import GraspEvaluator
    from crog_b200.model import CROG
    from oracle import grasp_tail_c as TC
    from oracle import warp_affine as WA

    Lw, B = 17, 2
    cfg = synth.default_cfg(Lw)
    model = CROG(cfg, precision="bf16")
    model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed"))
    model = model.cuda()
    img, word = synth.make_inputs(B, Lw)
    gt, cnt = synth.make_gt_rects(B, 64, seed=4)
    _, mat_inv = W.get_transform_mat(ori, (416, 416), inverse=True)
    rng = np.random.default_rng(1)
    tgt416 = (rng.random((B, 416, 416)) > 0.6).astype(np.float32)
    ev = GraspEvaluator(model)
    res = ev.step_original(img.cuda(), word.cuda(), torch.from_numpy(gt.copy()).cuda(), torch.from_numpy(cnt).cuda(),
                           mat_inv, ori, mask_target=torch.from_numpy(tgt416).cuda())
    post = res["post"].cpu().numpy()      # [5,B,416,416] as produced by the GPU
    inv = res["maps"].cpu().numpy()       # [5,B,h,w]
    for b in range(B):
        for p in range(5):
            assert np.array_equal(inv[p, b], WA.warp_affine_cubic_f32(post[p, b], mat_inv, (ori[1], ori[0]), 0.0)), (p, b)
        t = WA.warp_affine_cubic_f32(tgt416[b], mat_inv, (ori[1], ori[0]), 0.0)
        assert float(res["iou"][b]) == WA.mask_iou(inv[0, b], t)
    g_ref, n_ref, j_ref, _ = TC.tail_batch(inv[1], inv[2], inv[3], inv[4], gt, cnt)
    assert np.array_equal(res["n_peaks"].cpu().numpy(), n_ref)
    assert np.array_equal(res["j_flags"].cpu().numpy(), j_ref)
    for b in range(B):
        k = int(n_ref[b])
        assert np.array_equal(res["grasps"].cpu().numpy()[b, :k, :4], g_ref[b, :k, :4])
    s = ev.summary()
    assert s["n"] == B and 0.0 <= s["IoU"] <= 1.0 and set(s["Pr"]) == {"Pr@50", "Pr@60", "Pr@70", "Pr@80", "Pr@90"}
