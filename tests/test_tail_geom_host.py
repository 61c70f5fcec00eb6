"""Host-compiled check of crog_b200/csrc/tail_geom.h — the exact integer rasterisation the
CUDA Jaccard kernel uses — against the oracle (no GPU needed: the header is __host__ __device__)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import grasp_tail as T

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def geom():
    so = os.path.join(HERE, "host", "_build", "libtail_geom_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    src = os.path.join(HERE, "host", "tail_geom_host.cpp")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return C.CDLL(so)


def _counts(lib, p, g, mode):
    pa = np.asarray(p[:5], np.float64); ga = np.asarray(g[:5], np.float64)
    i, u, f = C.c_int(), C.c_int(), C.c_int()
    lib.tg_host_counts(pa.ctypes.data_as(C.POINTER(C.c_double)), ga.ctypes.data_as(C.POINTER(C.c_double)), mode,
                       C.byref(i), C.byref(u), C.byref(f))
    return i.value, u.value, f.value


def _oracle_counts(p, g):
    a = T.rect_pixels(g); b = T.rect_pixels(p)
    return len(a & b), len(a | b)


def test_box_points_host(geom):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    out = np.zeros(8, np.float32)
    for _ in range(2000):
        cx, cy = rng.uniform(-50, 700, 2); w, h = rng.uniform(0, 150, 2); a = rng.uniform(-200, 200)
        geom.tg_host_box_points(C.c_float(cx), C.c_float(cy), C.c_float(w), C.c_float(h), C.c_float(a),
                                out.ctypes.data_as(C.POINTER(C.c_float)))
        assert (out.reshape(4, 2) == cv2.boxPoints(((cx, cy), (w, h), a))).all()


def test_row_mask_rasterisation_matches_oracle(geom):
    rng = np.random.default_rng(7)
    n_fast = 0
    for it in range(400):
        if it % 4 == 0:  # typical grasp rectangles
            p = [rng.uniform(20, 460), rng.uniform(20, 460), rng.uniform(0, 110), 20, rng.uniform(-90, 90)]
        elif it % 4 == 1:  # near / across the raster border, incl. x >= 480
            p = [rng.uniform(-30, 520), rng.uniform(-30, 520), rng.uniform(0, 100), 20, rng.uniform(-90, 90)]
        elif it % 4 == 2:  # thin / degenerate after truncation
            p = [rng.uniform(20, 460), rng.uniform(20, 460), rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(-180, 180)]
        else:  # axis aligned, integer-ish corners (edge / vertex rules)
            p = [float(rng.integers(30, 450)), float(rng.integers(30, 450)), float(rng.integers(0, 60) * 2), 20.0,
                 float(rng.choice([0, 90, -90, 45, 180]))]
        g = [p[0] + rng.uniform(-25, 25), p[1] + rng.uniform(-25, 25), rng.uniform(0, 100), 20, p[4] + rng.uniform(-40, 40)]
        want = _oracle_counts(p, g)
        i0, u0, fast = _counts(geom, p, g, 0)
        i1, u1, _ = _counts(geom, p, g, 1)
        n_fast += fast
        assert (i0, u0) == want, (p, g)
        assert (i1, u1) == want, (p, g)
    assert n_fast > 300


def test_big_rectangles_take_slow_path(geom):
    p = [240, 240, 400, 300, 17]
    g = [250, 230, 380, 20, 10]
    want = _oracle_counts(p, g)
    i0, u0, fast = _counts(geom, p, g, 0)
    assert not fast and (i0, u0) == want
