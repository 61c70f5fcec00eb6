"""CPU: the warp oracle (oracle/warp_affine.py) against the cv2-generated golden vectors (tests/golden/warp_cases.npz,
written by oracle/make_golden_warp.py from the real cv2.warpAffine + the reference's torch normalisation) and, where
cv2 is importable, against cv2 itself on fresh random cases."""
import hashlib
import os

import numpy as np
import pytest

from oracle import warp_affine as WA


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "warp_cases.npz"))


def test_f32_small_cases_bit_exact(g):
    for i in range(int(g["n_f32"])):
        got = WA.warp_affine_cubic_f32(g[f"f32_{i}_src"], g[f"f32_{i}_M"], g[f"f32_{i}_out"].shape[::-1], float(g[f"f32_{i}_bv"]))
        assert np.array_equal(got, g[f"f32_{i}_out"]), i


def test_u8_preprocess_small_cases_bit_exact(g):
    for i in range(int(g["n_u8"])):
        got = WA.preprocess_image(g[f"u8_{i}_img"], g[f"u8_{i}_M"], tuple(int(v) for v in g[f"u8_{i}_size"]))
        assert np.array_equal(got, g[f"u8_{i}_out"]), i


def test_letterbox_full_size(g):
    rng = np.random.default_rng(int(g["lb_seed"]))
    maps = rng.random((5, 416, 416), dtype=np.float32)
    inv = np.stack([WA.warp_affine_cubic_f32(m, g["lb_mat_inv"], (640, 480), 0.0) for m in maps])
    assert _sha(inv) == str(g["lb_inv_sha"])
    assert np.array_equal(inv[:, ::16, ::16], g["lb_inv_sample"])
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    pre = WA.preprocess_image(img, g["lb_mat"], (416, 416))
    assert _sha(pre) == str(g["lb_pre_sha"])
    tgt416 = (rng.random((416, 416)) > 0.7).astype(np.float32)
    tgt = WA.warp_affine_cubic_f32(tgt416, g["lb_mat_inv"], (640, 480), 0.0)
    assert WA.mask_iou(inv[0], tgt) == float(g["lb_iou"])


def test_tables():
    t = WA.cubic_tab_1d()
    assert t.shape == (32, 4) and np.array_equal(t[0], np.array([0, 1, 0, 0], np.float32))
    ti = WA.cubic_tab_2d_i16()
    assert (ti.reshape(32, 32, 16).sum(-1) == WA.COEF_SCALE).all()  # initInterTab2D forces every kernel to sum to 2^15


def test_against_cv2_random():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(12):
        Hs, Ws, h, w = [int(v) for v in rng.integers(8, 120, 4)]
        ang, s = rng.uniform(-3.14, 3.14), rng.uniform(0.3, 3.0)
        M = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-60, 60)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-60, 60)]])
        src = rng.standard_normal((Hs, Ws)).astype(np.float32)
        bv = float(rng.choice([0.0, 1.5]))
        assert np.array_equal(cv2.warpAffine(src, M, (w, h), flags=cv2.INTER_CUBIC, borderValue=bv), WA.warp_affine_cubic_f32(src, M, (w, h), bv))
        img = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
        b3 = [123.4, 7.5, 250.5]
        assert np.array_equal(cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_CUBIC, borderValue=b3), WA.warp_affine_cubic_u8(img, M, (w, h), b3))
