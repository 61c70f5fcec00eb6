"""GPU parity of the grasp-decode + Jaccard tail against GOLDENS OF THE REFERENCE ITSELF.

tests/golden/tail_cases.npz holds what the unmodified reference ``utils/grasp_eval.py:289-374`` returned when executed by
oracle/make_golden_tail.py (real cv2.boxPoints; the two scikit-image calls served by scipy / OpenCV code, see
oracle/skimage_literal.py).  Bars: peak indices, x / y / width*100 / 20, pixel counts, IoU floats, J@1 / J@K decisions and the
in-place target edit bit-exact; the angle within 4 float32 ulp of the reference's libm value and within
1 ulp of the oracle's.  All calls go through the C ABI.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from crog_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tail_cases.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _detect_vs_gold(q, s, c, w, K, peaks_g, n_g, gref, gorc):
    from crog_b200.utils import grasp_eval as GE

    peaks, n, grasps = GE.detect_grasps_batched(*[torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in (q, s, c, w)], K)
    torch.cuda.synchronize()
    peaks, n, grasps = peaks.cpu().numpy(), n.cpu().numpy(), grasps.cpu().numpy()
    assert np.array_equal(n, n_g)
    assert np.array_equal(peaks, peaks_g)                      # (row, col), -1 padded: bit-exact vs the reference run
    for b in range(q.shape[0]):
        k = int(n_g[b])
        got = grasps[b, :k]
        assert np.array_equal(got[:, :4], gorc[b, :k, :4])     # NumPy-1.24 float64 rows (the pinned environment)
        assert np.array_equal(got[:, :2], gref[b, :k, :2]) and (got[:, 3] == 20).all()
        assert np.array_equal(got[:, 2].astype(np.float32), gref[b, :k, 2].astype(np.float32))  # width*100, exact
        ulp = np.spacing(np.abs(gorc[b, :k, 4]).astype(np.float32)).astype(np.float64)
        assert (np.abs(got[:, 4] - gorc[b, :k, 4]) <= ulp + 1e-12).all()
        assert (np.abs(got[:, 4] - gref[b, :k, 4]) <= 4 * ulp + 1e-12).all()


@pytest.mark.parametrize("K", [1, 5, 9])
def test_detect_small_maps_vs_reference_goldens(gold, K):
    _detect_vs_gold(gold["small_q"], gold["small_s"], gold["small_c"], gold["small_w"], K, gold[f"small_peaks_k{K}"],
                    gold[f"small_n_k{K}"], gold[f"small_gref_k{K}"], gold[f"small_gorc_k{K}"])


@pytest.mark.parametrize("kind", ["blobs", "stress"])
def test_detect_config5_maps_vs_reference_goldens(gold, kind):
    q, s, c, w = synth.make_tail_maps(int(gold[f"{kind}_n_maps"]), kind, seed=int(gold[f"{kind}_seed"]), size=416)
    assert _sha(q) + _sha(s) + _sha(c) + _sha(w) == str(gold[f"{kind}_sha"]), "synthetic generator drifted"
    _detect_vs_gold(q, s, c, w, 5, gold[f"{kind}_peaks"], gold[f"{kind}_n"], gold[f"{kind}_gref"], gold[f"{kind}_gorc"])
    # K = 1 is the first row of the K = 5 decode
    pk1 = np.full((q.shape[0], 1, 2), -1, np.int32)
    n1 = np.minimum(gold[f"{kind}_n"], 1).astype(np.int32)
    pk1[:, 0] = gold[f"{kind}_peaks"][:, 0]
    _detect_vs_gold(q, s, c, w, 1, pk1, n1, gold[f"{kind}_gref"], gold[f"{kind}_gorc"])


def test_iou_pairs_vs_reference_goldens(gold):
    """1600 (prediction, ground truth) pairs: x >= 480 quirk, canvas edges, unedited GT (w > 100, h != 20), degenerate
    rectangles, both angle-gate branches.  One pair per batch row (K = M = 1), targets NOT edited (calculate_iou)."""
    from crog_b200.utils import grasp_eval as GE

    P, G = gold["iou_p"], gold["iou_g"]
    n = len(P)
    g = torch.from_numpy(P.reshape(n, 1, 5).copy()).to(DEV)
    t = torch.from_numpy(G.reshape(n, 1, 6).copy()).to(DEV)
    cnt = torch.ones(n, dtype=torch.int32, device=DEV)
    _, inter, uni = GE.jacquard_batched(g, None, t, cnt, want_counts=True, edit_gt=False)
    torch.cuda.synchronize()
    inter, uni = inter.cpu().numpy()[:, 0, 0].astype(np.int64), uni.cpu().numpy()[:, 0, 0].astype(np.int64)
    assert np.array_equal(inter, gold["iou_inter"]) and np.array_equal(uni, gold["iou_union"])
    iou = np.where(uni > 0, inter / np.maximum(uni, 1), 0.0)
    assert np.array_equal(iou, gold["iou_ref"])                # the float calculate_iou of the reference returned
    assert np.array_equal(t.cpu().numpy().reshape(n, 6), G)    # calculate_iou never edits its arguments
    # reference-signature wrapper on a sample of the pairs
    for i in range(0, n, 97):
        assert GE.calculate_iou(list(P[i]), list(G[i])) == gold["iou_ref"][i]


def test_jaccard_cases_vs_reference_goldens(gold):
    from crog_b200.utils import grasp_eval as GE

    preds, npred, gt, cnt = gold["j_preds"], gold["j_npred"], gold["j_gt"], gold["j_cnt"]
    B = len(preds)
    gt_dev = torch.from_numpy(gt.copy()).to(DEV)
    counters = torch.zeros(4, dtype=torch.int64, device=DEV)
    flags = GE.jacquard_batched(torch.from_numpy(preds.copy()).to(DEV), torch.from_numpy(npred.copy()).to(DEV), gt_dev,
                                torch.from_numpy(cnt.copy()).to(DEV), counters=counters)
    torch.cuda.synchronize()
    flags = flags.cpu().numpy()
    assert np.array_equal(flags[:, 0], gold["j_at1"]) and np.array_equal(flags[:, 1], gold["j_atk"])
    after = gt_dev.cpu().numpy()
    for b in range(B):
        assert np.array_equal(after[b, :cnt[b]], gold["j_gt_after"][b, :cnt[b]]), b   # grasp_eval.py:367-368
    assert counters.cpu().tolist() == [int(gold["j_at1"].sum()), B, int(gold["j_atk"].sum()), B]
    # reference-signature wrappers: J flag, max IoU and the in-place edit of a float64 / float32 ndarray
    for b in range(0, B, 5):
        m = int(cnt[b])
        tg = gt[b, :m].copy()
        assert GE.calculate_jacquard_index([list(x) for x in preds[b, :npred[b]]], tg) == gold["j_atk"][b]
        assert np.array_equal(tg, gold["j_gt_after"][b, :m])
    b = 2
    tg = gold["j_gt_after"][b, :cnt[b]].copy()
    assert GE.calculate_max_iou([list(x) for x in preds[b, :npred[b]]], tg) == gold["j_max_iou"][b]
    tg32 = np.array([[200, 210, 150, 33, 5, 1], [100, 90, 40, 10, -20, 1]], np.float32)
    assert GE.calculate_jacquard_index(gold["j32_pred"].tolist(), tg32) == gold["j32_flag"]
    assert tg32.dtype == np.float32 and np.array_equal(tg32, gold["j32_gt_after"])
