"""Drop-in for the decode / Jaccard entry points of the reference's ``utils/grasp_eval.py``
(detect_grasps :289, calculate_iou :305, calculate_max_iou :350, calculate_jacquard_index :362),
running on the B200 through libcrog_b200.so.

The reference versions take one H x W numpy map / one rectangle at a time on the host.  The same
signatures are kept (numpy or torch inputs, Python lists / floats out) and each also has a
batched, device-resident form (``detect_grasps_batched``, ``jacquard_batched``) that the engine
uses so nothing leaves HBM between the model and the J@1 / J@5 counters.
There is no CPU implementation here: without a B200 every function raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .. import _lib as L

_ws_cache = {}


def _dev():
    if not torch.cuda.is_available():
        raise L.CrogError("crog_b200.utils.grasp_eval needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda_f32(a) -> torch.Tensor:
    t = torch.as_tensor(a)
    return t.to(_dev(), torch.float32).contiguous()


def _workspace(B, H, W, K, dev) -> torch.Tensor:
    key = (B, H, W, K, dev.index)
    ws = _ws_cache.get(key)
    if ws is None:
        n = L.load().crog_detect_workspace_bytes(B, H, W, K)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        if len(_ws_cache) > 8:
            _ws_cache.clear()
        _ws_cache[key] = ws
    return ws


def detect_grasps_batched(q: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, wid: torch.Tensor, num_grasps: int = 5,
                          threshold: float = 0.4) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """B x H x W float32 CUDA maps -> (peaks [B,K,2] int32 (row,col; -1 padded), n_peaks [B] int32,
    grasps [B,K,5] float64 rows [x, y, width*100, 20, angle_deg])."""
    lib = L.lib()
    assert q.is_cuda and q.dtype == torch.float32 and q.dim() == 3
    q, sin, cos, wid = [t.contiguous() for t in (q, sin, cos, wid)]
    B, H, W = q.shape
    K = int(num_grasps)
    peaks = torch.empty((B, K, 2), dtype=torch.int32, device=q.device)
    n = torch.empty((B,), dtype=torch.int32, device=q.device)
    grasps = torch.empty((B, K, 5), dtype=torch.float64, device=q.device)
    with torch.cuda.device(q.device):
        ws = _workspace(B, H, W, K, q.device)
        L.check(lib.crog_detect_grasps(q.data_ptr(), sin.data_ptr(), cos.data_ptr(), wid.data_ptr(), B, H, W, K,
                                       float(threshold), peaks.data_ptr(), n.data_ptr(), grasps.data_ptr(),
                                       ws.data_ptr(), L.stream_ptr()))
    return peaks, n, grasps


def angle_map(sin: torch.Tensor, cos: torch.Tensor) -> torch.Tensor:
    lib = L.lib()
    sin, cos = sin.contiguous(), cos.contiguous()
    out = torch.empty_like(sin)
    with torch.cuda.device(sin.device):
        L.check(lib.crog_angle_map(sin.data_ptr(), cos.data_ptr(), out.data_ptr(), sin.numel(), L.stream_ptr()))
    return out


def jacquard_batched(grasps: torch.Tensor, n_peaks: Optional[torch.Tensor], gt: torch.Tensor, gt_count: torch.Tensor,
                     counters: Optional[torch.Tensor] = None, want_counts: bool = False, edit_gt: bool = True):
    """grasps [B,K,5] f64, gt [B,M,6] f64 (edited in place like the reference when edit_gt), gt_count [B] i32 ->
    j_flags [B,2] int32 = (J@1, J@K) and optionally (inter, union) pixel counts [B,K,M] int32.
    ``counters`` (int64[4] on the device) accumulates correct@1,total@1,correct@K,total@K."""
    lib = L.lib()
    assert grasps.is_cuda and grasps.dtype == torch.float64 and gt.dtype == torch.float64 and gt.is_contiguous()
    grasps = grasps.contiguous()
    B, K, _ = grasps.shape
    M = gt.shape[1]
    flags = torch.empty((B, 2), dtype=torch.int32, device=grasps.device)
    inter = uni = None
    if want_counts:
        inter = torch.zeros((B, K, M), dtype=torch.int32, device=grasps.device)
        uni = torch.zeros((B, K, M), dtype=torch.int32, device=grasps.device)
    with torch.cuda.device(grasps.device):
        L.check(lib.crog_jaccard(grasps.data_ptr(), n_peaks.data_ptr() if n_peaks is not None else None, K, gt.data_ptr(),
                                 gt_count.data_ptr(), M, B, inter.data_ptr() if want_counts else None,
                                 uni.data_ptr() if want_counts else None, flags.data_ptr(),
                                 counters.data_ptr() if counters is not None else None, int(edit_gt), L.stream_ptr()))
    return (flags, inter, uni) if want_counts else flags


# ----------------------------------------------------------------------- reference-signature wrappers
def detect_grasps(grasp_quality_mask, grasp_sin_mask, grasp_cos_mask, grasp_wid_mask, num_grasps=5):
    """utils/grasp_eval.py:289-302: returns (list of [x, y, w, 20, angle_deg], angle map)."""
    is_np = not torch.is_tensor(grasp_quality_mask)
    q, s, c, w = [_as_cuda_f32(a)[None] for a in (grasp_quality_mask, grasp_sin_mask, grasp_cos_mask, grasp_wid_mask)]
    peaks, n, grasps = detect_grasps_batched(q, s, c, w, num_grasps)
    ang = angle_map(s[0], c[0])
    k = int(n[0].item())
    rows = grasps[0, :k].cpu().tolist()
    out = [[r[0], r[1], r[2], 20, r[4]] for r in rows]
    return out, (ang.cpu().numpy() if is_np else ang)


def calculate_iou(rect_p, rect_gt, shape=(480, 640), angle_threshold=30):
    """utils/grasp_eval.py:305-347.  Only the reference's default ``shape`` / ``angle_threshold`` exist on the device."""
    if tuple(shape) != (480, 640) or angle_threshold != 30:
        raise L.CrogError("calculate_iou: only shape=(480,640), angle_threshold=30 (the reference defaults) are implemented")
    dev = _dev()
    g = torch.tensor([[[float(v) for v in rect_p[:5]]]], dtype=torch.float64, device=dev)
    t = torch.tensor([[[float(v) for v in rect_gt[:5]] + [0.0]]], dtype=torch.float64, device=dev)
    cnt = torch.ones(1, dtype=torch.int32, device=dev)
    _, inter, uni = jacquard_batched(g, None, t, cnt, want_counts=True, edit_gt=False)
    i, u = int(inter[0, 0, 0].item()), int(uni[0, 0, 0].item())
    return 0 if u <= 0 else i / u


def calculate_max_iou(rects_p, rects_gt):
    """utils/grasp_eval.py:350-359."""
    best = 0
    for g in rects_gt:
        for p in rects_p:
            v = calculate_iou(p, g)
            if v > best:
                best = v
    return best


def calculate_jacquard_index(grasp_preds, grasp_targets, iou_threshold=0.25):
    """utils/grasp_eval.py:362-374, including the in-place edit of ``grasp_targets`` (h:=20, w clipped to [0,100])."""
    if iou_threshold != 0.25:
        raise L.CrogError("calculate_jacquard_index: only iou_threshold=0.25 (the reference default) is implemented")
    dev = _dev()
    preds = np.asarray(grasp_preds, dtype=np.float64).reshape(-1, 5)
    is_np = isinstance(grasp_targets, np.ndarray)
    tg = np.asarray(grasp_targets, dtype=np.float64)
    K = max(preds.shape[0], 1)
    g = torch.zeros((1, K, 5), dtype=torch.float64, device=dev)
    if preds.shape[0]:
        g[0, :preds.shape[0]] = torch.from_numpy(preds).to(dev)
    n = torch.tensor([preds.shape[0]], dtype=torch.int32, device=dev)
    t = torch.from_numpy(np.ascontiguousarray(tg[None, :, :6] if tg.shape[1] >= 6 else np.pad(tg, ((0, 0), (0, 6 - tg.shape[1])))[None])).to(dev)
    cnt = torch.tensor([tg.shape[0]], dtype=torch.int32, device=dev)
    flags = jacquard_batched(g, n, t, cnt)
    edited = t[0].cpu().numpy()
    if is_np and grasp_targets.dtype.kind == "f":
        grasp_targets[:, 2] = edited[:, 2]
        grasp_targets[:, 3] = edited[:, 3]
    return int(flags[0, 1].item())
