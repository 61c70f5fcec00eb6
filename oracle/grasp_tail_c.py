"""ORACLE: ctypes binding of oracle/grasp_tail.c (plain-C restatement of
utils/grasp_eval.py:289-374).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgrasp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "grasp_tail.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_peak_local_max.restype = C.c_int
        _lib.oracle_detect_grasps.restype = C.c_int
        _lib.oracle_iou_counts.restype = C.c_int
        _lib.oracle_jacquard.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def peak_local_max(img, thr=0.4, num_peaks=5):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros((num_peaks, 2), np.int32)
    n = lib().oracle_peak_local_max(_p(img, C.c_float), img.shape[0], img.shape[1], C.c_float(thr), num_peaks,
                                    _p(out, C.c_int))
    return out[:n].astype(np.int64)


def detect_grasps(q, s, c, w, num_grasps=5):
    q, s, c, w = [np.ascontiguousarray(a, np.float32) for a in (q, s, c, w)]
    out = np.zeros((num_grasps, 5), np.float64)
    rc = np.zeros((num_grasps, 2), np.int32)
    n = lib().oracle_detect_grasps(_p(q, C.c_float), _p(s, C.c_float), _p(c, C.c_float), _p(w, C.c_float),
                                   q.shape[0], q.shape[1], num_grasps, _p(out, C.c_double), _p(rc, C.c_int))
    return out[:n], rc[:n]


def box_points(cx, cy, w, h, ang):
    out = np.zeros(8, np.float32)
    lib().oracle_box_points(C.c_float(cx), C.c_float(cy), C.c_float(w), C.c_float(h), C.c_float(ang), _p(out, C.c_float))
    return out.reshape(4, 2)


def iou_counts(rect_p, rect_gt):
    p = np.ascontiguousarray(rect_p, np.float64)
    g = np.ascontiguousarray(rect_gt, np.float64)
    i, u = C.c_long(0), C.c_long(0)
    lib().oracle_iou_counts(_p(p, C.c_double), _p(g, C.c_double), C.byref(i), C.byref(u))
    return i.value, u.value


def jacquard(preds, gts):
    """gts [M,6] float64 is edited in place like the reference."""
    p = np.ascontiguousarray(preds, np.float64).reshape(-1, 5)
    assert gts.dtype == np.float64 and gts.flags.c_contiguous
    return lib().oracle_jacquard(_p(p, C.c_double), p.shape[0], _p(gts, C.c_double), gts.shape[0])


def tail_batch(q, s, c, w, gts, cnt):
    """Serial reference loop over a batch; returns grasps[B,5,5], n[B], j[B,2], counters[4]."""
    q, s, c, w = [np.ascontiguousarray(a, np.float32) for a in (q, s, c, w)]
    B, H, W = q.shape
    gts = np.ascontiguousarray(gts, np.float64).copy()
    cnt = np.ascontiguousarray(cnt, np.int32)
    grasps = np.zeros((B, 5, 5), np.float64); n = np.zeros(B, np.int32); j = np.zeros((B, 2), np.int32)
    counters = np.zeros(4, np.int64)
    lib().oracle_tail_batch(_p(q, C.c_float), _p(s, C.c_float), _p(c, C.c_float), _p(w, C.c_float), B, H, W,
                            _p(gts, C.c_double), _p(cnt, C.c_int), gts.shape[1], _p(grasps, C.c_double),
                            _p(n, C.c_int), _p(j, C.c_int), _p(counters, C.c_long))
    return grasps, n, j, counters
