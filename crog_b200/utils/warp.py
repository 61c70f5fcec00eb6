"""Device versions of the two ``cv2.warpAffine(..., INTER_CUBIC)`` steps that sit either side of the model in
the reference pipeline, bit-exact with OpenCV (kernels: crog_b200/csrc/warp.cu):

  * ``preprocess_images``  — utils/dataset.py:843-866: letterbox the uint8 RGB image to the network input size with
    the CLIP-mean border, then ``/255, -mean, /std`` into the float32 NCHW tensor ``CROG.forward`` takes;
  * ``warp_affine_cubic``  — engine/crog_engine.py:387-391,499-517: the inverse letterbox of the prediction (and
    target) maps back to the original image size before mask-IoU and ``detect_grasps``;
  * ``mask_iou``           — engine/crog_engine.py:500-501,515-518.

The 2x3 matrices are the ones the reference's dataset builds with ``cv2.getAffineTransform``
(``get_transform_mat``, utils/dataset.py:825-840); OpenCV inverts them in float64 inside ``warpAffine`` and
``invert_affine`` repeats that arithmetic on the host, so the kernels receive destination->source maps.
There is no CPU implementation here: without a B200 every function that touches pixels raises.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L

_pre_ws = {}
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def get_transform_mat(img_size: Tuple[int, int], input_size: Tuple[int, int] = (416, 416), inverse: bool = False):
    """utils/dataset.py:825-840 (same name, same return): (mat, mat_inv | None), float64 2x3."""
    ori_h, ori_w = img_size
    inp_h, inp_w = input_size
    scale = min(inp_h / ori_h, inp_w / ori_w)
    new_h, new_w = ori_h * scale, ori_w * scale
    bias_x, bias_y = (inp_w - new_w) / 2.0, (inp_h - new_h) / 2.0
    src = np.array([[0, 0], [ori_w, 0], [0, ori_h]], np.float32)
    dst = np.array([[bias_x, bias_y], [new_w + bias_x, bias_y], [bias_x, new_h + bias_y]], np.float32)
    try:
        import cv2  # the reference's own dependency; gives the identical LU-solved matrix

        mat = cv2.getAffineTransform(src, dst)
        mat_inv = cv2.getAffineTransform(dst, src) if inverse else None
    except ImportError:  # closed form of the same 3-point system (axis-aligned letterbox)
        def solve(a, b):
            A = np.zeros((6, 6)); rhs = np.zeros(6)
            for i in range(3):
                A[i, 0:2], A[i, 2] = a[i], 1
                A[i + 3, 3:5], A[i + 3, 5] = a[i], 1
                rhs[i], rhs[i + 3] = b[i, 0], b[i, 1]
            return np.linalg.solve(A, rhs).reshape(2, 3)

        mat = solve(src.astype(np.float64), dst.astype(np.float64))
        mat_inv = solve(dst.astype(np.float64), src.astype(np.float64)) if inverse else None
    return mat, mat_inv


def invert_affine(M) -> np.ndarray:
    """What cv::warpAffine does to M when WARP_INVERSE_MAP is not set (float64) -> flat [6]."""
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


class DeviceAffine:
    """Affine matrices already inverted (invert_affine) and resident on the device as [B,6] float64: what the kernels
    consume.  Build once with ``device_affine(mats, B, dev)`` and pass it wherever ``mats`` is accepted, so a streaming
    loop does not pay a host inversion + pageable H2D copy (a stream synchronisation) per batch."""

    def __init__(self, t: torch.Tensor):
        self.t = t


def device_affine(mats, B: int, dev) -> DeviceAffine:
    return DeviceAffine(_minv_tensor(mats, B, torch.device(dev)))


def _minv_tensor(mats, B: int, dev) -> torch.Tensor:
    """One 2x3 matrix (shared) or B of them, as the caller would pass to cv2.warpAffine -> device [B,6] float64."""
    if isinstance(mats, DeviceAffine):
        if mats.t.shape != (B, 6) or mats.t.device != dev:
            raise RuntimeError(f"DeviceAffine of shape {tuple(mats.t.shape)} on {mats.t.device} does not fit batch {B} on {dev}")
        return mats.t
    if torch.is_tensor(mats):
        mats = mats.detach().cpu().numpy()
    m = np.asarray(mats, dtype=np.float64)
    if m.ndim == 2:
        m = np.broadcast_to(m, (B, 2, 3))
    if m.shape != (B, 2, 3):
        raise RuntimeError(f"expected {B} affine matrices of shape 2x3, got {m.shape}")
    inv = np.stack([invert_affine(x) for x in m])
    return torch.from_numpy(inv).to(dev)


def _dev_of(t: torch.Tensor):
    if not t.is_cuda:
        raise L.CrogError("crog_b200.utils.warp needs CUDA tensors (sm_100a); there is no CPU fallback")
    return t.device


def warp_affine_cubic(src: torch.Tensor, mats, dsize: Tuple[int, int], border_value: float = 0.0) -> torch.Tensor:
    """``cv2.warpAffine(src_i, mats_b, dsize, flags=cv2.INTER_CUBIC, borderValue=border_value)`` for every plane:
    src [NP,B,Hs,Ws] (or [B,Hs,Ws]) float32 CUDA, dsize = (w, h) as in OpenCV -> [NP,B,h,w] (or [B,h,w])."""
    lib = L.lib()
    dev = _dev_of(src)
    squeeze = src.dim() == 3
    s = (src[None] if squeeze else src).contiguous().float()
    NP, B, Hs, Ws = s.shape
    w, h = int(dsize[0]), int(dsize[1])
    minv = _minv_tensor(mats, B, dev)
    out = torch.empty((NP, B, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_warp_affine_cubic_f32(s.data_ptr(), NP, B, Hs, Ws, minv.data_ptr(), out.data_ptr(), h, w,
                                               float(border_value), L.stream_ptr()))
    return out[0] if squeeze else out


def preprocess_images(img_u8: torch.Tensor, mats, input_size: Tuple[int, int] = (416, 416),
                      mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD, out: torch.Tensor = None) -> torch.Tensor:
    """utils/dataset.py:843-866 for a batch of equally sized images: img_u8 [B,Ho,Wo,3] uint8 RGB CUDA, ``mats`` the
    forward letterbox matrices (``get_transform_mat(...)[0]``) -> [B,3,H,W] float32 normalised (written into ``out`` when
    given, e.g. ``CROG.input_buffer(B)`` so the forward reads it in place)."""
    import ctypes as C

    lib = L.lib()
    dev = _dev_of(img_u8)
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
        raise RuntimeError(f"expected uint8 images [B,H,W,3], got {img_u8.dtype} {tuple(img_u8.shape)}")
    img_u8 = img_u8.contiguous()
    B, Ho, Wo, _ = img_u8.shape
    Sh, Sw = int(input_size[0]), int(input_size[1])
    minv = _minv_tensor(mats, B, dev)
    if out is None:
        out = torch.empty((B, 3, Sh, Sw), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (B, 3, Sh, Sw) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise RuntimeError(f"out must be a contiguous float32 [{B},3,{Sh},{Sw}] tensor on {dev}")
    border = (C.c_double * 3)(*[float(np.float64(m) * 255) for m in CLIP_MEAN])  # dataset.py:860
    mean32 = torch.tensor(list(mean), dtype=torch.float32).numpy()
    std32 = torch.tensor(list(std), dtype=torch.float32).numpy()
    cm, cs = (C.c_float * 3)(*mean32.tolist()), (C.c_float * 3)(*std32.tolist())
    ws = _pre_ws.get(dev.index)
    if ws is None:  # OpenCV's 32 x 32 table of int16 bicubic weight sets; the call rebuilds it on the stream (ordered)
        ws = _pre_ws[dev.index] = torch.empty(int(L.load().crog_preprocess_workspace_bytes()), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_preprocess_u8(img_u8.data_ptr(), B, Ho, Wo, minv.data_ptr(), out.data_ptr(), Sh, Sw, border, cm, cs,
                                       ws.data_ptr(), L.stream_ptr()))
    return out


def mask_iou(pred: torch.Tensor, target: torch.Tensor, thr: float = 0.35) -> Tuple[torch.Tensor, torch.Tensor]:
    """pred, target [B,h,w] float32 CUDA -> (iou [B] float64 = inter / (union + 1e-6), counts [B,2] int64)."""
    lib = L.lib()
    dev = _dev_of(pred)
    pred, target = pred.contiguous().float(), target.contiguous().float()
    B = pred.shape[0]
    n = pred[0].numel() if B else 0
    counts = torch.empty((B, 2), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_mask_iou(pred.data_ptr(), target.data_ptr(), B, n, float(thr), counts.data_ptr(), L.stream_ptr()))
    iou = counts[:, 0].double() / (counts[:, 1].double() + 1e-6)
    return iou, counts
