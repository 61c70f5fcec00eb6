"""Pins the tail oracle (oracle/grasp_tail.py and oracle/grasp_tail.c) against every
known-answer vector available for the path, the installed
cv2 / scipy building blocks, and each other."""
import numpy as np
import pytest

from oracle import grasp_tail as T
from oracle import grasp_tail_c as TC
from crog_b200 import synth


def test_peak_local_max_docstring_vector():
    # skimage.feature.peak_local_max docstring (A.6): min_distance=2 keeps only [3,2]
    img = np.zeros((7, 7), np.float32)
    img[3, 4] = 1
    img[3, 2] = 1.5
    assert T.peak_local_max(img, 2, 0.4, 5).tolist() == [[3, 2]]
    assert TC.peak_local_max(img, 0.4, 5).tolist() == [[3, 2]]


def test_polygon_docstring_vector():
    # skimage.draw.polygon docstring (A.6): r=[1,2,8], c=[1,7,4] on a 10x10 image
    rr, cc = T.polygon([1, 2, 8], [1, 7, 4])
    img = np.zeros((10, 10), np.uint8)
    img[rr, cc] = 1
    want = np.zeros((10, 10), np.uint8)
    rows = {1: [1], 2: range(2, 8), 3: range(2, 7), 4: range(3, 7), 5: range(3, 6), 6: [4, 5], 7: [4], 8: [4]}
    for r, cs in rows.items():
        for c in cs:
            want[r, c] = 1
    assert (img == want).all()


def test_box_points_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(3000):
        cx, cy = rng.uniform(-50, 700, 2)
        w, h = rng.uniform(0, 150, 2)
        a = rng.uniform(-200, 200)
        want = cv2.boxPoints(((cx, cy), (w, h), a))
        assert (T.box_points(cx, cy, w, h, a) == want).all()
        assert (TC.box_points(cx, cy, w, h, a) == want).all()


def test_max_filter_matches_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(1)
    img = rng.random((37, 53), dtype=np.float32)
    want = ndi.maximum_filter(img, footprint=np.ones((5, 5)), mode="nearest")
    assert (T.max_filter5(img) == want).all()


def test_peak_rules_ties_border_trivial():
    # constant image -> trivial -> no peaks
    assert len(T.peak_local_max(np.full((20, 20), 0.7, np.float32))) == 0
    assert len(TC.peak_local_max(np.full((20, 20), 0.7, np.float32))) == 0
    # plateau: equal neighbours, greedy keeps the first in row-major order, rejects 8-neighbours,
    # keeps points at Chebyshev distance exactly 2
    img = np.zeros((12, 12), np.float32)
    img[5, 4:8] = 0.9
    got = T.peak_local_max(img, num_peaks=5).tolist()
    assert got == [[5, 4], [5, 6]]
    assert TC.peak_local_max(img, 0.4, 5).tolist() == got
    # border of width 2 excluded; threshold is strict in float32
    img = np.zeros((12, 12), np.float32)
    img[1, 5] = 1.0; img[6, 6] = np.float32(0.4); img[8, 3] = np.nextafter(np.float32(0.4), np.float32(1))
    assert T.peak_local_max(img, num_peaks=5).tolist() == [[8, 3]]
    assert TC.peak_local_max(img, 0.4, 5).tolist() == [[8, 3]]


@pytest.mark.parametrize("kind", ["blobs", "stress"])
def test_py_and_c_oracles_agree_on_maps(kind):
    q, s, c, w = synth.make_tail_maps(4 if kind == "blobs" else 20, kind, seed=11, size=96)
    for i in range(q.shape[0]):
        for K in (1, 5):
            g_py, ang = T.detect_grasps(q[i], s[i], c[i], w[i], K)
            g_c, rc = TC.detect_grasps(q[i], s[i], c[i], w[i], K)
            assert len(g_py) == len(g_c)
            if len(g_py):
                assert np.array_equal(np.asarray(g_py, np.float64), g_c), (kind, i, K)


def test_iou_py_and_c_agree_and_quirks():
    rng = np.random.default_rng(3)
    for _ in range(150):
        p = [rng.uniform(0, 520), rng.uniform(0, 500), rng.uniform(0, 110), 20, rng.uniform(-90, 90)]
        g = [p[0] + rng.uniform(-30, 30), p[1] + rng.uniform(-30, 30), rng.uniform(0, 100), 20,
             p[4] + rng.uniform(-50, 50), 1.0]
        assert T.iou_counts(p, g) == TC.iou_counts(p, g)
    # x >= 480 is silently dropped (A.4 quirk): a rectangle fully right of x=480 has no pixels
    assert T.iou_counts([560, 200, 60, 20, 0], [560, 200, 60, 20, 0, 1]) == (0, 0)
    assert TC.iou_counts([560, 200, 60, 20, 0], [560, 200, 60, 20, 0, 1]) == (0, 0)
    # identical rectangles -> IoU 1; angle gate -> 0
    i, u = T.iou_counts([200, 200, 60, 20, 10], [200, 200, 60, 20, 10, 1])
    assert i == u and i > 0
    assert T.calculate_iou([200, 200, 60, 20, 45], [200, 200, 60, 20, -5, 1]) == 0  # |d|=50, |s|=40
    # gate passes through the "sum" branch (|a+b| <= 30)
    assert T.calculate_iou([200, 200, 60, 20, 80], [200, 200, 60, 20, -85, 1]) > 0


def test_jacquard_inplace_edit_and_empty_preds():
    g = np.array([[200., 200., 150., 33., 5., 1.]])
    p = [[200., 200., 100., 20, 5.]]
    assert T.calculate_jacquard_index(p, g) == 1
    assert g[0, 3] == 20 and g[0, 2] == 100  # in-place overwrite/clip (grasp_eval.py:367-368)
    g2 = np.array([[200., 200., 150., 33., 5., 1.]])
    assert TC.jacquard(np.asarray(p, np.float64), g2) == 1 and g2[0, 2] == 100
    assert T.calculate_jacquard_index(np.zeros((0, 5)), np.array([[200., 200., 50., 20., 5., 1.]])) == 0


def test_tail_batch_c_matches_python():
    q, s, c, w = synth.make_tail_maps(3, "blobs", seed=21, size=128)
    gt, cnt = synth.make_gt_rects(3, 8, seed=5, size=128)
    grasps, n, j, counters = TC.tail_batch(q, s, c, w, gt, cnt)
    for b in range(3):
        g5, _ = T.detect_grasps(q[b], s[b], c[b], w[b], 5)
        g1, _ = T.detect_grasps(q[b], s[b], c[b], w[b], 1)
        assert n[b] == len(g5)
        j1 = T.calculate_jacquard_index(g1, gt[b, :cnt[b]].copy()) if len(g1) else 0
        j5 = T.calculate_jacquard_index(g5, gt[b, :cnt[b]].copy()) if len(g5) else 0
        assert (j[b, 0], j[b, 1]) == (j1, j5)
    assert counters[1] == 3 and counters[3] == 3


# ------------------------------------------------------------------ pins added in round 2
# tests/golden/tail_cases.npz is written by oracle/make_golden_tail.py, which EXECUTES the reference's own
# utils/grasp_eval.py:289-374 (real cv2.boxPoints; skimage stubbed by oracle/skimage_literal.py, i.e. by scipy / OpenCV code)
# and asserts reference == oracle before writing.  These tests re-check the oracle against those vectors anywhere.
import hashlib
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tail_cases.npz")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _check_rows(g_ref, g_orc, rows, k_n):
    """rows: oracle rows (float64); g_ref: what the unmodified reference returned under NumPy 2 (float32 w / angle)."""
    assert len(rows) == k_n
    for k in range(k_n):
        assert np.array_equal(np.asarray(rows[k], np.float64), g_orc[k])
        assert rows[k][0] == g_ref[k, 0] and rows[k][1] == g_ref[k, 1] and rows[k][3] == 20
        assert np.float32(rows[k][2]) == np.float32(g_ref[k, 2])
        assert abs(rows[k][4] - g_ref[k, 4]) <= 4 * float(np.spacing(np.float32(abs(rows[k][4])))) + 1e-12


@pytest.mark.parametrize("K", [1, 5, 9])
def test_golden_detect_small_maps(gold, K):
    q, s, c, w = gold["small_q"], gold["small_s"], gold["small_c"], gold["small_w"]
    for i in range(q.shape[0]):
        n = int(gold[f"small_n_k{K}"][i])
        rows, _ = T.detect_grasps(q[i], s[i], c[i], w[i], K)
        _check_rows(gold[f"small_gref_k{K}"][i], gold[f"small_gorc_k{K}"][i], rows, n)
        rows_c, rc = TC.detect_grasps(q[i], s[i], c[i], w[i], K)
        assert np.array_equal(rc, gold[f"small_peaks_k{K}"][i, :n])
        assert np.array_equal(rows_c, gold[f"small_gorc_k{K}"][i, :n])


@pytest.mark.parametrize("kind", ["blobs", "stress"])
def test_golden_detect_config5_maps(gold, kind):
    q, s, c, w = synth.make_tail_maps(int(gold[f"{kind}_n_maps"]), kind, seed=int(gold[f"{kind}_seed"]), size=416)
    assert _sha(q) + _sha(s) + _sha(c) + _sha(w) == str(gold[f"{kind}_sha"]), "synthetic generator drifted"
    for i in range(q.shape[0]):
        n = int(gold[f"{kind}_n"][i])
        rows_c, rc = TC.detect_grasps(q[i], s[i], c[i], w[i], 5)
        assert np.array_equal(rc, gold[f"{kind}_peaks"][i, :n]) and len(rc) == n
        _check_rows(gold[f"{kind}_gref"][i], gold[f"{kind}_gorc"][i], [list(r) for r in rows_c], n)
    for i in range(0, q.shape[0], 7):  # the numpy restatement is slower: a sample
        rows, _ = T.detect_grasps(q[i], s[i], c[i], w[i], 5)
        _check_rows(gold[f"{kind}_gref"][i], gold[f"{kind}_gorc"][i], rows, int(gold[f"{kind}_n"][i]))


def test_golden_iou_pairs(gold):
    P, G = gold["iou_p"], gold["iou_g"]
    for i in range(len(P)):
        ii, uu = TC.iou_counts(P[i], G[i])
        assert (ii, uu) == (gold["iou_inter"][i], gold["iou_union"][i]), i
        assert (0 if uu <= 0 else ii / uu) == gold["iou_ref"][i], i   # the float the reference returned
    for i in range(0, len(P), 9):
        assert T.iou_counts(P[i], G[i]) == (gold["iou_inter"][i], gold["iou_union"][i]), i


def test_golden_jaccard_cases(gold):
    preds, npred, gt, cnt = gold["j_preds"], gold["j_npred"], gold["j_gt"], gold["j_cnt"]
    for b in range(len(preds)):
        m = int(cnt[b])
        tg = np.ascontiguousarray(gt[b, :m].copy())
        assert TC.jacquard(preds[b, :npred[b]], tg) == gold["j_atk"][b], b
        assert np.array_equal(tg, gold["j_gt_after"][b, :m])
        tg1 = np.ascontiguousarray(gt[b, :m].copy())
        assert TC.jacquard(preds[b, :min(npred[b], 1)], tg1) == gold["j_at1"][b], b
    for b in range(0, len(preds), 6):
        m = int(cnt[b])
        tg = gt[b, :m].copy()
        assert T.calculate_jacquard_index(preds[b, :npred[b]].reshape(-1, 5), tg) == gold["j_atk"][b]
        assert float(T.calculate_max_iou(preds[b, :npred[b]], tg)) == gold["j_max_iou"][b]
    tg32 = np.array([[200, 210, 150, 33, 5, 1], [100, 90, 40, 10, -20, 1]], np.float32)
    assert T.calculate_jacquard_index(gold["j32_pred"], tg32) == gold["j32_flag"] and np.array_equal(tg32, gold["j32_gt_after"])


def test_polygon_equals_cv2_point_polygon_test():
    """oracle polygon (O'Rourke crossing test restated) == {p : cv2.pointPolygonTest >= 0} on 10 000 random truncated
    cv2.boxPoints quadrilaterals of non-zero area (both exact integer tests; independent installed code)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import skimage_literal as SL

    rng = np.random.default_rng(5)
    n = deg = 0
    while n < 10000:
        cx, cy = rng.uniform(-20, 500), rng.uniform(-20, 660)
        big = n % 8 == 0
        w = rng.uniform(0, 110) if big else rng.uniform(0, 24)
        h = rng.choice([20.0, rng.uniform(0, 40)]) if big else rng.uniform(0, 12)
        a = rng.choice([0.0, 90.0, 45.0, rng.uniform(-180, 180)])
        box = cv2.boxPoints(((cx, cy), (w, h), a)).astype(np.int64)
        x, y = box[:, 0], box[:, 1]
        if sum(int(x[i]) * int(y[(i + 1) % 4]) - int(x[(i + 1) % 4]) * int(y[i]) for i in range(4)) == 0:
            deg += 1
            continue
        n += 1
        rr, cc = T.polygon(x, y, (480, 640))
        r2, c2 = SL.polygon_cv2(x, y, (480, 640))
        assert np.array_equal(rr, r2) and np.array_equal(cc, c2), box.tolist()
    assert deg > 50  # zero-area quads exist in the distribution; they are covered by the reference-run goldens


def test_zero_area_quads_keep_only_vertices():
    """A doubled segment has no interior and O'Rourke's test reports only its end points (the published skimage rule);
    OpenCV would call the whole segment 'on the edge'.  The C and numpy restatements agree with each other."""
    rr, cc = T.polygon([203, 203, 181, 181], [112, 112, 46, 46], (480, 640))
    assert sorted(zip(rr.tolist(), cc.tolist())) == [(181, 46), (203, 112)]
    assert T.iou_counts([200, 200, 0.4, 20, 0], [200, 200, 0.4, 20, 0, 1]) == TC.iou_counts([200, 200, 0.4, 20, 0], [200, 200, 0.4, 20, 0, 1])


def test_greedy_spacing_equals_literal_ensure_spacing():
    """The oracle's single greedy pass == skimage's batched cKDTree ensure_spacing written out literally
    (oracle/skimage_literal.py), on maps with thousands of tied candidates (several 50/+100/+200 batches, the max_out
    break, np.delete) and for several K."""
    pytest.importorskip("scipy.spatial")
    from oracle import skimage_literal as SL

    rng = np.random.default_rng(8)
    maps = []
    for levels in (2, 3, 8, 64):
        maps.append(np.floor(rng.random((90, 120)) * levels).astype(np.float32) / levels * 0.6 + 0.41)
    m = np.full((70, 70), 0.5, np.float32); m[::3, ::3] = 0.8          # lattice of equal peaks at distance 3
    maps.append(m)
    m = np.zeros((70, 70), np.float32); m[10:60, 10:60] = 0.9         # one big plateau
    maps.append(m)
    maps.append(rng.random((150, 150), dtype=np.float32))               # iid: ~900 candidates
    q, _, _, _ = synth.make_tail_maps(2, "stress", seed=3, size=200)
    maps += [q[0], q[1]]
    for i, m in enumerate(maps):
        for K in (1, 5, 40, 400):
            got = T.peak_local_max(m, 2, 0.4, K)
            want = SL.peak_local_max_literal(m, 2, 0.4, K)
            assert np.array_equal(got, want), (i, K)
    # num_peaks=inf: no early break at all
    m = maps[0]
    got = T.peak_local_max(m, 2, 0.4, 10 ** 9)
    want = SL.peak_local_max_literal(m, 2, 0.4, np.inf)
    assert np.array_equal(got, want)
