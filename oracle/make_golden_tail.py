"""ORACLE tooling: pin oracle/grasp_tail.py (+ grasp_tail.c) against the reference's OWN tail code and write
tests/golden/tail_cases.npz.

Run in the build container only (needs /root/reference and OpenCV):
    python oracle/make_golden_tail.py

The reference's ``utils/grasp_eval.py`` is imported UNMODIFIED (same harness as oracle/make_golden_ssg.py: scikit-image
and matplotlib, absent offline, are satisfied by stub modules; ``np.int0`` is restored) and its own lines 289-374 are
executed — ``detect_grasps``, ``calculate_iou`` (real ``cv2.boxPoints``, the x/y transposition, the ``rr < 640`` /
``cc < 480`` filters, the ``area[cc, rr]`` canvas), ``calculate_max_iou`` and ``calculate_jacquard_index`` (in-place
target edit) — next to the oracle on the same inputs; equality is asserted before anything is written.

The two scikit-image functions the reference calls are stubbed by code that does NOT come from the oracle's loops
wherever installed third-party code can stand in (oracle/skimage_literal.py): ``peak_local_max`` = real scipy maximum
filter + the literal cKDTree ``ensure_spacing``; ``polygon`` = cv2.pointPolygonTest >= 0 for non-degenerate quadrilaterals
(zero-area ones fall back to the oracle's point-in-polygon routine: OpenCV counts the interior of a doubled segment as
"on the edge", O'Rourke's crossing test as published in skimage does not — the one place the two differ).

NumPy note: the reference environment pins NumPy 1.24.3, whose value-based promotion makes ``width*100`` and
``angle/pi*180`` float64; under the NumPy 2 installed here the unmodified lines compute them in float32.  ``width*100`` is
exact in float64, so float32(oracle) must equal the reference's value bit for bit; the angle is compared to 4 float32 ulp
(libm/SIMD atan2f is not reproducible).  IoU / Jaccard take float64 rows and are promotion-free:
they are compared exactly.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from crog_b200 import synth  # noqa: E402
from oracle import grasp_tail as T  # noqa: E402
from oracle import grasp_tail_c as TC  # noqa: E402
from oracle import skimage_literal as SL  # noqa: E402


def _polygon_stub(r, c, shape=None):
    r = np.asarray(r, np.int64); c = np.asarray(c, np.int64)
    n = len(r)
    area2 = sum(int(c[i]) * int(r[(i + 1) % n]) - int(c[(i + 1) % n]) * int(r[i]) for i in range(n))
    if area2 == 0:
        return T.polygon(r, c, shape)
    return SL.polygon_cv2(r, c, shape)


def import_reference_grasp_eval():
    sk = types.ModuleType("skimage")
    sk_draw, sk_filters, sk_feature = types.ModuleType("skimage.draw"), types.ModuleType("skimage.filters"), types.ModuleType("skimage.feature")
    sk_draw.polygon = _polygon_stub
    sk_filters.gaussian = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not on the tail path"))
    sk_feature.peak_local_max = SL.peak_local_max_literal
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    for name, m in (("skimage", sk), ("skimage.draw", sk_draw), ("skimage.filters", sk_filters), ("skimage.feature", sk_feature),
                    ("matplotlib", mpl), ("matplotlib.pyplot", plt)):
        sys.modules[name] = m
    if not hasattr(np, "int0"):
        np.int0 = np.int64
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    try:
        import utils.grasp_eval as ref_ge  # reference, unmodified
    finally:
        sys.path.remove(REF)
    assert os.path.realpath(ref_ge.__file__).startswith(REF)
    return ref_ge


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --------------------------------------------------------------------------------------------- detect_grasps
def small_maps():
    """Twelve 48x48 maps stored verbatim: constants, plateaus, ties, border and strict-threshold cases, noise."""
    S = 48
    rng = np.random.default_rng(101)
    q = np.zeros((12, S, S), np.float32)
    q[0] = 0.7                                                        # trivial image -> no peaks
    q[1, 20:30, 10:40] = 0.9                                          # plateau: stable order + spacing
    q[2, 1, 5] = 1.0; q[2, 30, 31] = np.float32(0.4); q[2, 40, 41] = np.nextafter(np.float32(0.4), np.float32(1))
    q[3] = np.float32(0.5); q[3, 10, 10] = 0.49                       # almost-constant map
    q[4] = np.floor(rng.random((S, S)) * 4).astype(np.float32) / 4    # heavy ties
    q[5, 2, 2] = 0.8; q[5, S - 3, S - 3] = 0.8; q[5, 2, S - 3] = 0.8  # first / last interior pixels
    q[6] = rng.random((S, S), dtype=np.float32)                       # iid noise
    q[7] = np.floor(rng.random((S, S)) * 16).astype(np.float32) / 16
    q[8, 10, 10] = 0.9; q[8, 10, 12] = 0.9; q[8, 11, 11] = 0.95       # neighbours at Chebyshev distance 1 and 2
    q[9] = np.linspace(0.3, 0.95, S, dtype=np.float32)[None, :].repeat(S, 0)  # ramp: one candidate column
    q[10] = 0.39                                                      # everything below the threshold
    q[11] = rng.random((S, S), dtype=np.float32) * 0.2 + 0.41         # everything above it
    s = rng.normal(size=q.shape).astype(np.float32); c = rng.normal(size=q.shape).astype(np.float32)
    w = rng.random(q.shape).astype(np.float32)
    return q, s, c, w


def detect_block(ref_ge, q, s, c, w, K):
    n_maps = q.shape[0]
    peaks = np.full((n_maps, K, 2), -1, np.int32)
    npk = np.zeros(n_maps, np.int32)
    g_ref = np.zeros((n_maps, K, 5), np.float64)
    g_orc = np.zeros((n_maps, K, 5), np.float64)
    for i in range(n_maps):
        r_list, r_ang = ref_ge.detect_grasps(q[i], s[i], c[i], w[i], K)     # the reference's own lines 289-302
        o_list, o_ang = T.detect_grasps(q[i], s[i], c[i], w[i], K)
        c_rows, c_rc = TC.detect_grasps(q[i], s[i], c[i], w[i], K)
        assert len(r_list) == len(o_list) == len(c_rows), (i, len(r_list), len(o_list))
        ra = np.asarray(r_ang, np.float32)
        assert np.all(np.abs(ra - o_ang) <= 4 * np.spacing(np.maximum(np.abs(ra), np.abs(o_ang)).astype(np.float32)))
        npk[i] = len(r_list)
        for k, (r, o) in enumerate(zip(r_list, o_list)):
            assert r[0] == o[0] and r[1] == o[1] and r[3] == o[3] == 20, (i, k, r, o)
            assert np.float32(o[2]) == np.float32(r[2]), (i, k, r[2], o[2])                 # width*100: exact
            assert abs(float(r[4]) - o[4]) <= 4 * float(np.spacing(np.float32(abs(o[4])))) + 1e-12  # angle: <= 4 ulp
            assert np.array_equal(np.asarray(o, np.float64), c_rows[k])
            peaks[i, k] = (int(r[1]), int(r[0]))
            g_ref[i, k] = [float(v) for v in r]
            g_orc[i, k] = o
    return peaks, npk, g_ref, g_orc


# --------------------------------------------------------------------------------------------- IoU pairs
def iou_pairs(n: int, seed: int):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 5)); G = np.zeros((n, 6))
    for i in range(n):
        kind = i % 8
        cx, cy = rng.uniform(20, 460), rng.uniform(20, 600)
        th = rng.uniform(-90, 90)
        w, h = rng.uniform(5, 100), 20.0
        if kind == 1:   # x >= 480: the rr < 640 / cc < 480 filters drop columns silently (A.4 quirk)
            cx = rng.uniform(440, 620)
        elif kind == 2:  # crosses the top / left canvas edge (negative truncated vertices, minr = max(0, .))
            cx, cy = rng.uniform(-15, 25), rng.uniform(-15, 25)
        elif kind == 3:  # GT as it is BEFORE calculate_jacquard_index edits it: w > 100, h != 20
            w, h = rng.uniform(100, 160), rng.uniform(5, 45)
        elif kind == 4:  # thin / tiny rectangles (degenerate after truncation)
            w, h = rng.uniform(0, 3), rng.choice([0.0, 0.5, 20.0])
        elif kind == 5:  # axis-aligned and 45 degrees
            th = float(rng.choice([0, 90, -90, 45, -45]))
        elif kind == 6:  # bottom edge: y up to 640 is inside the transposed canvas test, beyond is clipped by polygon()
            cy = rng.uniform(600, 660)
        G[i] = [cx, cy, w, h, th, 1.0]
        dth = rng.uniform(-28, 28) if i % 3 else rng.uniform(-80, 80)  # both sides of the 30-degree gate
        if i % 11 == 0:  # the |a + b| <= 30 branch of the gate
            dth = -2 * th + rng.uniform(-25, 25)
        P[i] = [cx + rng.uniform(-14, 14), cy + rng.uniform(-14, 14), rng.uniform(0, 105), 20.0, th + dth]
        if i % 17 == 0:
            P[i] = [G[i, 0], G[i, 1], G[i, 2], G[i, 3], G[i, 4]]       # identical -> IoU 1 (or 0/0 if no pixels)
    return P, G


def iou_block(ref_ge, P, G):
    n = len(P)
    iou = np.zeros(n, np.float64); inter = np.zeros(n, np.int64); union = np.zeros(n, np.int64)
    for i in range(n):
        r = ref_ge.calculate_iou(list(P[i]), list(G[i]))                 # the reference's own lines 305-347
        ii, uu = T.iou_counts(P[i], G[i])
        assert (ii, uu) == TC.iou_counts(P[i], G[i]), i
        want = 0 if uu <= 0 else ii / uu
        assert float(r) == float(want), (i, P[i], G[i], r, want)
        iou[i], inter[i], union[i] = float(r), ii, uu
    return iou, inter, union


# --------------------------------------------------------------------------------------------- Jaccard sets
def jaccard_cases(n: int, seed: int, K: int = 5, M: int = 64):
    rng = np.random.default_rng(seed)
    gt, cnt = synth.make_gt_rects(n, M, seed=seed + 1)
    gt[:, :, 0] += rng.uniform(0, 120, gt.shape[:2])                    # some GT beyond x = 480
    preds = np.zeros((n, K, 5)); npred = rng.integers(0, K + 1, n).astype(np.int32)
    npred[:4] = [0, 1, K, K]
    for b in range(n):
        for k in range(K):
            m = int(rng.integers(0, cnt[b]))
            if k % 2 == 0 or b % 3 == 0:
                preds[b, k] = [gt[b, m, 0] + rng.uniform(-10, 10), gt[b, m, 1] + rng.uniform(-10, 10), rng.uniform(0, 105), 20,
                               gt[b, m, 4] + rng.uniform(-33, 33)]
            else:
                preds[b, k] = [rng.uniform(0, 500), rng.uniform(0, 500), rng.uniform(0, 105), 20, rng.uniform(-90, 90)]
    return preds, npred, gt, cnt


def jaccard_block(ref_ge, preds, npred, gt, cnt):
    n = len(preds)
    j1 = np.zeros(n, np.int32); jk = np.zeros(n, np.int32); miou = np.zeros(n, np.float64)
    gt_after = gt.copy()
    for b in range(n):
        m = int(cnt[b])
        p = preds[b, :npred[b]]
        tg_r = gt[b, :m].copy(); tg_o = gt[b, :m].copy(); tg_c = gt[b, :m].copy()
        r = ref_ge.calculate_jacquard_index([list(x) for x in p], tg_r)  # the reference's own lines 362-374
        o = T.calculate_jacquard_index(p.reshape(-1, 5), tg_o)
        cc = TC.jacquard(p, tg_c)
        assert r == o == cc, (b, r, o, cc)
        assert np.array_equal(tg_r, tg_o) and np.array_equal(tg_r, tg_c)  # identical in-place edits
        assert (tg_r[:, 3] == 20).all() and tg_r[:, 2].max() <= 100
        mi = ref_ge.calculate_max_iou([list(x) for x in p], tg_r)
        assert float(mi) == float(T.calculate_max_iou(p, tg_o)), b
        assert (mi > 0.25) == bool(r)
        r1 = ref_ge.calculate_jacquard_index([list(x) for x in p[:1]], gt[b, :m].copy())
        assert r1 == T.calculate_jacquard_index(p[:1].reshape(-1, 5), gt[b, :m].copy())
        j1[b], jk[b], miou[b] = r1, r, float(mi)
        gt_after[b, :m] = tg_r
    return j1, jk, miou, gt_after


def main():
    ref_ge = import_reference_grasp_eval()
    out = {}
    # ---- detect: verbatim small maps, K = 1, 5, 9
    q, s, c, w = small_maps()
    out.update(small_q=q, small_s=s, small_c=c, small_w=w)
    for K in (1, 5, 9):
        pk, n, gr, go = detect_block(ref_ge, q, s, c, w, K)
        out.update({f"small_peaks_k{K}": pk, f"small_n_k{K}": n, f"small_gref_k{K}": gr, f"small_gorc_k{K}": go})
        print(f"small maps K={K}: peaks per map {n.tolist()}")
    # ---- detect: config-5 maps at full size, regenerated from seeds by the tests (hash-checked)
    for kind, n_maps, seed in (("blobs", 8, 40), ("stress", 40, 41)):
        q, s, c, w = synth.make_tail_maps(n_maps, kind, seed=seed, size=416)
        pk, n, gr, go = detect_block(ref_ge, q, s, c, w, 5)
        pk1, n1, _, _ = detect_block(ref_ge, q, s, c, w, 1)
        assert np.array_equal(pk1[:, 0], pk[:, 0]) and np.array_equal(n1, np.minimum(n, 1))  # J@1 = first row of the K=5 decode
        out.update({f"{kind}_n_maps": np.int64(n_maps), f"{kind}_seed": np.int64(seed), f"{kind}_sha": np.array(sha(q) + sha(s) + sha(c) + sha(w)),
                    f"{kind}_peaks": pk, f"{kind}_n": n, f"{kind}_gref": gr, f"{kind}_gorc": go})
        print(f"{kind}: {n_maps} maps 416x416, peaks per map {n.tolist()}")
    # ---- calculate_iou
    P, G = iou_pairs(1600, 7)
    iou, inter, union = iou_block(ref_ge, P, G)
    out.update(iou_p=P, iou_g=G, iou_ref=iou, iou_inter=inter, iou_union=union)
    print(f"iou pairs: {len(P)}, gated/empty {(union == 0).sum()}, overlapping {(inter > 0).sum()}, IoU>0.25 {(iou > 0.25).sum()}, "
          f"x>=480 touched {(G[:, 0] > 470).sum()}")
    assert (inter > 0).sum() > 400 and (union == 0).sum() > 100
    # ---- calculate_max_iou / calculate_jacquard_index
    preds, npred, gt, cnt = jaccard_cases(48, 23)
    j1, jk, miou, gt_after = jaccard_block(ref_ge, preds, npred, gt, cnt)
    out.update(j_preds=preds, j_npred=npred, j_gt=gt, j_cnt=cnt, j_at1=j1, j_atk=jk, j_max_iou=miou, j_gt_after=gt_after)
    print(f"jaccard cases: {len(preds)}, J@1 hits {int(j1.sum())}, J@5 hits {int(jk.sum())}")
    assert 5 <= jk.sum() <= len(preds) - 5, "both outcomes should be well represented"
    # float32 targets: the in-place edit happens in the array's own dtype; integer-dtype targets are not a valid input of the
    # reference (cv2.boxPoints rejects numpy integer scalars), so they are not part of the contract
    tg32 = np.array([[200, 210, 150, 33, 5, 1], [100, 90, 40, 10, -20, 1]], np.float32)
    r = ref_ge.calculate_jacquard_index([[201.0, 209.0, 90.0, 20, 7.0]], tg32)
    tg_o = np.array([[200, 210, 150, 33, 5, 1], [100, 90, 40, 10, -20, 1]], np.float32)
    assert r == T.calculate_jacquard_index(np.array([[201.0, 209.0, 90.0, 20, 7.0]]), tg_o) and np.array_equal(tg32, tg_o)
    assert tg32.dtype == np.float32 and tg32[0, 2] == 100 and tg32[0, 3] == 20
    out.update(j32_pred=np.array([[201.0, 209.0, 90.0, 20, 7.0]]), j32_gt_after=tg32, j32_flag=np.int32(r))
    path = os.path.join(ROOT, "tests", "golden", "tail_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
