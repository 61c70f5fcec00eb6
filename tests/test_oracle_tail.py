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
