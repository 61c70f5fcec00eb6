"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of ``cv2.warpAffine(..., flags=cv2.INTER_CUBIC,
borderMode=BORDER_CONSTANT)`` as the reference calls it:

  * engine/crog_engine.py:387-391,499-517 — the inverse letterbox of the five float32 prediction maps
    (416x416 -> ori_size, borderValue=0) before mask-IoU and detect_grasps;
  * utils/dataset.py:843-866            — the forward letterbox of the uint8 RGB image
    (ori_size -> 416x416, borderValue = CLIP mean * 255) before normalisation.

The arithmetic lives in a third-party dependency that is not under /root/reference: opencv-python==4.7.0.72
(environment.yml:61).  Its published algorithm (modules/imgproc/src/imgwarp.cpp: warpAffine, WarpAffineInvoker,
initInterTab2D, remapBicubic) is restated here:

  1. M is inverted in float64 (no WARP_INVERSE_MAP flag is passed);
  2. source coordinates are 22.10 fixed point: adelta[x] = rint(M0*x*1024), X0 = rint((M1*y + M2)*1024) + 16,
     X = (X0 + adelta[x]) >> 5  -> integer pixel X >> 5 (saturated to int16) and a 1/32-pixel phase X & 31;
  3. weights come from a 32 x 32 table of outer products of the 1-D cubic (A = -0.75) coefficients, float32
     for float images, int16 (scale 2^15, corrected to sum exactly to 2^15) for uint8 images;
  4. the 4x4 footprint is summed row group by row group in float32 in the interior, tap by tap against the
     constant border value where the footprint leaves the image, and is the border value when fully outside.

PINNED: tests/test_oracle_warp.py compares every function here bit-for-bit with the cv2 installed in the build
container (4.13; the algorithm is unchanged since 2.x) on random matrices, sizes and the letterbox cases, and
tests/golden/warp_*.npz holds cv2-generated vectors (oracle/make_golden_warp.py) for the GPU box.
"""
from __future__ import annotations

import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
ROUND_DELTA = AB_SCALE // INTER_TAB_SIZE // 2
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


def invert_affine(M) -> np.ndarray:
    """cv::warpAffine's in-place inversion of the 2x3 matrix (float64), imgwarp.cpp."""
    M = np.array(M, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def _sat_int(v: np.ndarray) -> np.ndarray:
    """saturate_cast<int>(double): round half to even, clamp to int32."""
    r = np.rint(v)
    return np.clip(r, -2147483648.0, 2147483647.0).astype(np.int64)


def source_coords(Minv: np.ndarray, w: int, h: int):
    """-> (sx, sy int32 [h,w] saturated to int16 range, ax, ay phases in [0,32))."""
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    adelta = _sat_int(Minv[0] * x * AB_SCALE)
    bdelta = _sat_int(Minv[3] * x * AB_SCALE)
    X0 = _sat_int((Minv[1] * y + Minv[2]) * AB_SCALE) + ROUND_DELTA
    Y0 = _sat_int((Minv[4] * y + Minv[5]) * AB_SCALE) + ROUND_DELTA
    # the C code adds in 32-bit ints (wraps); coordinates that large are saturated anyway
    X = ((X0[:, None] + adelta[None, :]).astype(np.int32)) >> (AB_BITS - INTER_BITS)
    Y = ((Y0[:, None] + bdelta[None, :]).astype(np.int32)) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767).astype(np.int32)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767).astype(np.int32)
    return sx, sy, (X & (INTER_TAB_SIZE - 1)).astype(np.int32), (Y & (INTER_TAB_SIZE - 1)).astype(np.int32)


def cubic_tab_1d() -> np.ndarray:
    """initInterTab1D(INTER_CUBIC): [32, 4] float32, every operation rounded to float32 like the C code."""
    f = np.float32
    A = f(-0.75)
    x = (np.arange(INTER_TAB_SIZE, dtype=np.float32) * f(1.0 / INTER_TAB_SIZE)).astype(np.float32)
    t = np.zeros((INTER_TAB_SIZE, 4), np.float32)
    x1 = x + f(1)
    t[:, 0] = ((A * x1 - f(5) * A) * x1 + f(8) * A) * x1 - f(4) * A
    t[:, 1] = ((A + f(2)) * x - (A + f(3))) * x * x + f(1)
    xm = f(1) - x
    t[:, 2] = ((A + f(2)) * xm - (A + f(3))) * xm * xm + f(1)
    t[:, 3] = f(1) - t[:, 0] - t[:, 1] - t[:, 2]
    return t


_TAB_F = None
_TAB_I = None


def cubic_tab_2d_f32() -> np.ndarray:
    """BicubicTab_f: [32(ay), 32(ax), 4(row), 4(col)] float32 = vy * vx."""
    global _TAB_F
    if _TAB_F is None:
        t = cubic_tab_1d()
        _TAB_F = (t[:, None, :, None] * t[None, :, None, :]).astype(np.float32)
    return _TAB_F


def cubic_tab_2d_i16() -> np.ndarray:
    """BicubicTab_i: saturate_cast<short>(v * 32768) with the sum forced to 32768 (initInterTab2D)."""
    global _TAB_I
    if _TAB_I is None:
        v = cubic_tab_2d_f32()
        it = np.clip(np.rint((v * np.float32(COEF_SCALE)).astype(np.float32).astype(np.float64)), -32768, 32767).astype(np.int32)
        for i in range(INTER_TAB_SIZE):
            for j in range(INTER_TAB_SIZE):
                tab = it[i, j]
                isum = int(tab.sum())
                if isum != COEF_SCALE:
                    diff = isum - COEF_SCALE
                    Mk1 = Mk2 = mk1 = mk2 = 2
                    for k1 in range(2, 4):
                        for k2 in range(2, 4):
                            if tab[k1, k2] < tab[mk1, mk2]:
                                mk1, mk2 = k1, k2
                            elif tab[k1, k2] > tab[Mk1, Mk2]:
                                Mk1, Mk2 = k1, k2
                    if diff < 0:
                        tab[Mk1, Mk2] = np.int16(tab[Mk1, Mk2] - diff)
                    else:
                        tab[mk1, mk2] = np.int16(tab[mk1, mk2] - diff)
        _TAB_I = it
    return _TAB_I


def warp_affine_cubic_f32(src: np.ndarray, M, dsize, border_value: float = 0.0, inverse_map: bool = False) -> np.ndarray:
    """cv2.warpAffine(src[H,W] float32, M, (w,h), flags=INTER_CUBIC, borderValue=border_value) -> [h,w] float32."""
    src = np.ascontiguousarray(src, dtype=np.float32)
    assert src.ndim == 2
    w, h = int(dsize[0]), int(dsize[1])
    Minv = np.array(M, np.float64).reshape(6) if inverse_map else invert_affine(M)
    sx, sy, ax, ay = source_coords(Minv, w, h)
    sx = sx - 1
    sy = sy - 1
    H, W = src.shape
    wt = cubic_tab_2d_f32()[ay, ax]  # [h, w, 4, 4]
    cv = np.float32(border_value)
    out = np.full((h, w), cv, np.float32)
    inside = (sx >= 0) & (sx < max(W - 3, 0)) & (sy >= 0) & (sy < max(H - 3, 0))
    outside = (sx >= W) | (sx + 4 <= 0) | (sy >= H) | (sy + 4 <= 0)
    # ---- interior: sum = r0; sum += r1; ... with r_i = ((S0*w0 + S1*w1) + S2*w2) + S3*w3, all float32
    iy, ix = np.nonzero(inside)
    if iy.size:
        x0, y0, wv = sx[iy, ix], sy[iy, ix], wt[iy, ix]
        total = None
        for i in range(4):
            r = None
            for j in range(4):
                p = src[y0 + i, x0 + j] * wv[:, i, j]
                r = p if r is None else (r + p).astype(np.float32)
            total = r if total is None else (total + r).astype(np.float32)
        out[iy, ix] = total
    # ---- border: sum = cv; for every tap inside the image, in row-major order: sum += (S - cv) * w
    by, bx = np.nonzero(~inside & ~outside)
    if by.size:
        x0, y0, wv = sx[by, bx], sy[by, bx], wt[by, bx]
        total = np.full(by.shape, cv * np.float32(1), np.float32)
        for i in range(4):
            yy = y0 + i
            oky = (yy >= 0) & (yy < H)
            for j in range(4):
                xx = x0 + j
                ok = oky & (xx >= 0) & (xx < W)
                v = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
                upd = (total + ((v - cv).astype(np.float32) * wv[:, i, j]).astype(np.float32)).astype(np.float32)
                total = np.where(ok, upd, total)
        out[by, bx] = total
    return out


def warp_affine_cubic_u8(src: np.ndarray, M, dsize, border_value=(0, 0, 0), inverse_map: bool = False) -> np.ndarray:
    """cv2.warpAffine(src[H,W,C] uint8, M, (w,h), flags=INTER_CUBIC, borderValue=...) -> [h,w,C] uint8 (fixed point)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    H, W, C = src.shape
    w, h = int(dsize[0]), int(dsize[1])
    bv = np.atleast_1d(np.asarray(border_value, dtype=np.float64))
    bv = np.concatenate([bv, np.zeros(max(0, 4 - bv.size))])[:max(C, 1)] if bv.size < C else bv[:C]
    cval = np.clip(np.rint(bv), 0, 255).astype(np.int64)  # saturate_cast<uchar>(double)
    Minv = np.array(M, np.float64).reshape(6) if inverse_map else invert_affine(M)
    sx, sy, ax, ay = source_coords(Minv, w, h)
    sx = sx - 1
    sy = sy - 1
    wt = cubic_tab_2d_i16()[ay, ax].astype(np.int64)  # [h, w, 4, 4]
    outside = (sx >= W) | (sx + 4 <= 0) | (sy >= H) | (sy + 4 <= 0)
    acc = np.broadcast_to(cval * COEF_SCALE, (h, w, C)).copy()
    s64 = src.astype(np.int64)
    for i in range(4):
        yy = sy + i
        oky = (yy >= 0) & (yy < H)
        for j in range(4):
            xx = sx + j
            ok = oky & (xx >= 0) & (xx < W)
            v = s64[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]  # [h, w, C]
            acc += np.where(ok[:, :, None], (v - cval) * wt[:, :, i, j][:, :, None], 0)
    res = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255)
    res = np.where(outside[:, :, None], cval, res).astype(np.uint8)
    return res[:, :, 0] if squeeze else res


def letterbox_mats(ori_size, input_size):
    """utils/dataset.py:825-840 get_transform_mat(img_size, inverse=True): the three point pairs the reference hands to
    cv2.getAffineTransform.  Returns (src_pts, dst_pts) float32; the matrices themselves come from cv2 (host side)."""
    ori_h, ori_w = ori_size
    inp_h, inp_w = input_size
    scale = min(inp_h / ori_h, inp_w / ori_w)
    new_h, new_w = ori_h * scale, ori_w * scale
    bias_x, bias_y = (inp_w - new_w) / 2.0, (inp_h - new_h) / 2.0
    src = np.array([[0, 0], [ori_w, 0], [0, ori_h]], np.float32)
    dst = np.array([[bias_x, bias_y], [new_w + bias_x, bias_y], [bias_x, new_h + bias_y]], np.float32)
    return src, dst


def preprocess_image(img_u8: np.ndarray, mat, input_size=(416, 416)) -> np.ndarray:
    """utils/dataset.py:843-866: letterbox (cubic, CLIP-mean border) -> CHW float32 -> /255, -mean, /std."""
    mean = np.array([0.48145466, 0.4578275, 0.40821073], np.float32).reshape(3, 1, 1)
    std = np.array([0.26862954, 0.26130258, 0.27577711], np.float32).reshape(3, 1, 1)
    border = [0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255]
    w = warp_affine_cubic_u8(img_u8, mat, (input_size[1], input_size[0]), border)
    x = w.transpose(2, 0, 1).astype(np.float32)
    x = (x / np.float32(255.0)).astype(np.float32)
    x = (x - mean).astype(np.float32)
    return (x / std).astype(np.float32)


def mask_iou(pred: np.ndarray, target: np.ndarray, thr: float = 0.35) -> float:
    """engine/crog_engine.py:500-501,515-518: (pred > 0.35) vs the warped target mask, float64 IoU."""
    p = pred > thr
    t = target.astype(bool) if target.dtype != np.bool_ else target
    inter = np.logical_and(p, t)
    union = np.logical_or(p, t)
    return float(np.sum(inter) / (np.sum(union) + 1e-6))
