"""ORACLE tooling: pin oracle/warp_affine.py against the real OpenCV + torch ops the reference calls and write
tests/golden/warp_cases.npz (run in the build container; needs cv2):

    python oracle/make_golden_warp.py

For every case the expected output is produced by ``cv2.warpAffine(..., flags=cv2.INTER_CUBIC, borderValue=...)``
(engine/crog_engine.py:387-391; utils/dataset.py:857-861) and, for the pre-processing cases, by the reference's own
torch arithmetic ``img.float().div_(255.).sub_(mean).div_(std)`` (utils/dataset.py:863-866) on CPU; the oracle must
agree bit for bit before anything is written.  Small cases store the full output; the full-size letterbox cases store
inputs by seed and outputs as a strided sample plus a SHA-256 of the bytes.
"""
from __future__ import annotations

import hashlib
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import warp_affine as WA  # noqa: E402

MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073]).reshape(3, 1, 1)
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711]).reshape(3, 1, 1)
BORDER = [0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255]


def rand_mat(rng):
    ang, s = rng.uniform(-np.pi, np.pi), rng.uniform(0.4, 2.5)
    return np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-30, 30)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-30, 30)]])


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_preprocess(img_u8, mat, size):
    w = cv2.warpAffine(img_u8, mat, (size[1], size[0]), flags=cv2.INTER_CUBIC, borderValue=BORDER)
    t = torch.from_numpy(w.transpose((2, 0, 1))).float()
    t.div_(255.).sub_(MEAN).div_(STD)
    return t.numpy()


def main():
    rng = np.random.default_rng(11)
    out = {"cv2_version": cv2.__version__}
    # ---- small float32 cases: random similarity transforms, three border values
    n_small = 6
    for i in range(n_small):
        Hs, Ws, h, w = [int(v) for v in rng.integers(24, 96, 4)]
        M = rand_mat(rng)
        src = rng.standard_normal((Hs, Ws)).astype(np.float32)
        bv = [0.0, 0.0, 0.5, -1.25, 0.0, 2.0][i]
        ref = cv2.warpAffine(src, M, (w, h), flags=cv2.INTER_CUBIC, borderValue=bv)
        got = WA.warp_affine_cubic_f32(src, M, (w, h), bv)
        assert np.array_equal(ref, got), f"f32 case {i}"
        out.update({f"f32_{i}_src": src, f"f32_{i}_M": M, f"f32_{i}_bv": np.float32(bv), f"f32_{i}_out": ref})
    out["n_f32"] = n_small
    # ---- small uint8 pre-processing cases
    n_u8 = 4
    for i in range(n_u8):
        Ho, Wo = [int(v) for v in rng.integers(30, 90, 2)]
        S = (int(rng.integers(32, 80)),) * 2
        img = rng.integers(0, 256, (Ho, Wo, 3), dtype=np.uint8)
        mat = cv2.getAffineTransform(*WA.letterbox_mats((Ho, Wo), S)) if i < 2 else rand_mat(rng)
        ref = ref_preprocess(img, mat, S)
        got = WA.preprocess_image(img, mat, S)
        assert np.array_equal(ref, got), f"u8 case {i}: {np.abs(ref - got).max()}"
        out.update({f"u8_{i}_img": img, f"u8_{i}_M": mat, f"u8_{i}_size": np.array(S), f"u8_{i}_out": ref})
    out["n_u8"] = n_u8
    # ---- full-size letterbox cases (OCID-VLG: 480x640 <-> 416x416), inputs by seed
    src_pts, dst_pts = WA.letterbox_mats((480, 640), (416, 416))
    mat, mat_inv = cv2.getAffineTransform(src_pts, dst_pts), cv2.getAffineTransform(dst_pts, src_pts)
    g = np.random.default_rng(12)
    maps = g.random((5, 416, 416), dtype=np.float32)
    inv = np.stack([cv2.warpAffine(m, mat_inv, (640, 480), flags=cv2.INTER_CUBIC, borderValue=0.) for m in maps])
    got = np.stack([WA.warp_affine_cubic_f32(m, mat_inv, (640, 480), 0.0) for m in maps])
    assert np.array_equal(inv, got), "letterbox inverse"
    img = g.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    pre = ref_preprocess(img, mat, (416, 416))
    assert np.array_equal(pre, WA.preprocess_image(img, mat, (416, 416))), "letterbox forward"
    out.update({"lb_mat": mat, "lb_mat_inv": mat_inv, "lb_seed": 12, "lb_inv_sha": sha(inv), "lb_inv_sample": inv[:, ::16, ::16].copy(),
                "lb_pre_sha": sha(pre), "lb_pre_sample": pre[:, ::16, ::16].copy()})
    # mask IoU of crog_engine.py:500-518 on the warped mask vs a warped synthetic target
    tgt416 = (g.random((416, 416)) > 0.7).astype(np.float32)
    tgt = cv2.warpAffine(tgt416, mat_inv, (640, 480), flags=cv2.INTER_CUBIC, borderValue=0.)
    p = inv[0] > 0.35
    iou = np.sum(np.logical_and(p, tgt)) / (np.sum(np.logical_or(p, tgt)) + 1e-6)
    assert iou == WA.mask_iou(inv[0], tgt)
    out["lb_iou"] = np.float64(iou)
    path = os.path.join(ROOT, "tests", "golden", "warp_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
