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

import os
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _lib as L

_ws_cache = {}
MAX_PREDS = 32  # predictions per sample one crog_jaccard call scores (MAXK in csrc/tail.cu)


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


_side_streams = {}
_pools = {}


def _stream_pool(dev, n: int):
    key = (dev.index, n)
    if key not in _pools:
        _pools[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _pools[key]


def decode_and_score_batched(q: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, wid: torch.Tensor, gt: torch.Tensor,
                             gt_count: torch.Tensor, num_grasps: int = 5, counters: Optional[torch.Tensor] = None,
                             chunks: int = 4, threshold: float = 0.4, edit_gt: bool = True):
    """detect_grasps (grasp_eval.py:289-302) followed by calculate_jacquard_index (:362-374) over a batch, issued as
    `chunks` sub-batches on two streams: the HBM-bound peak scan of sub-batch i+1 runs while the integer rasterisation
    of sub-batch i (ALU / latency bound, ~3 KB per sample) occupies the otherwise idle issue slots.  Same kernels, same
    bytes, same results as detect_grasps_batched + jacquard_batched; returns (peaks, n_peaks, grasps, j_flags)."""
    lib = L.lib()
    assert q.is_cuda and q.dtype == torch.float32 and q.dim() == 3
    assert gt.is_cuda and gt.dtype == torch.float64 and gt.is_contiguous() and gt_count.dtype == torch.int32
    q, sin, cos, wid = [t.contiguous() for t in (q, sin, cos, wid)]
    B, H, W = q.shape
    K, M = int(num_grasps), gt.shape[1]
    dev = q.device
    peaks = torch.empty((B, K, 2), dtype=torch.int32, device=dev)
    n = torch.empty((B,), dtype=torch.int32, device=dev)
    grasps = torch.empty((B, K, 5), dtype=torch.float64, device=dev)
    flags = torch.empty((B, 2), dtype=torch.int32, device=dev)
    chunks = max(1, min(int(chunks), B))
    per = (B + chunks - 1) // chunks
    cptr = counters.data_ptr() if counters is not None else None
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        side = _side_streams.get(dev.index)
        if side is None:
            side = _side_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
        ws = _workspace(per, H, W, K, dev)  # the detect kernels of successive sub-batches are ordered on `main`
        for b0 in range(0, B, per):
            nb = min(per, B - b0)
            L.check(lib.crog_detect_grasps(q[b0:].data_ptr(), sin[b0:].data_ptr(), cos[b0:].data_ptr(), wid[b0:].data_ptr(), nb, H, W,
                                           K, float(threshold), peaks[b0:].data_ptr(), n[b0:].data_ptr(), grasps[b0:].data_ptr(),
                                           ws.data_ptr(), main.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            L.check(lib.crog_jaccard(grasps[b0:].data_ptr(), n[b0:].data_ptr(), K, gt[b0:].data_ptr(), gt_count[b0:].data_ptr(), M, nb,
                                     None, None, flags[b0:].data_ptr(), cptr, int(edit_gt), side.cuda_stream))
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)
    return peaks, n, grasps, flags


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
    t = torch.from_numpy(np.ascontiguousarray(tg[None, :, :6] if tg.shape[1] >= 6 else np.pad(tg, ((0, 0), (0, 6 - tg.shape[1])))[None])).to(dev)
    cnt = torch.tensor([tg.shape[0]], dtype=torch.int32, device=dev)
    hit = 0
    # the device kernel scores at most MAX_PREDS predictions per call; longer lists go in chunks (max over chunks = max over
    # all pairs; the target edit is idempotent)
    for p0 in range(0, max(preds.shape[0], 1), MAX_PREDS):
        chunk = preds[p0:p0 + MAX_PREDS]
        K = max(chunk.shape[0], 1)
        g = torch.zeros((1, K, 5), dtype=torch.float64, device=dev)
        if chunk.shape[0]:
            g[0, :chunk.shape[0]] = torch.from_numpy(chunk).to(dev)
        n = torch.tensor([chunk.shape[0]], dtype=torch.int32, device=dev)
        flags = jacquard_batched(g, n, t, cnt)
        hit |= int(flags[0, 1].item())
    if is_np:  # the reference edits any ndarray in place, in the array's own dtype
        edited = t[0].cpu().numpy()
        grasp_targets[:, 2] = edited[:, 2]
        grasp_targets[:, 3] = edited[:, 3]
    return hit


# ======================================================================= SSG post-processing (BASELINE config 4)
_anchor_cache = {}


def _anchors_dev(anchors, dev) -> torch.Tensor:
    if torch.is_tensor(anchors):
        return anchors.to(dev, torch.float32).reshape(-1, 4).contiguous()
    key = (id(anchors), len(anchors), dev.index)
    t = _anchor_cache.get(key)
    if t is None:
        t = torch.tensor(anchors, dtype=torch.float32, device=dev).reshape(-1, 4).contiguous()
        if len(_anchor_cache) > 8:
            _anchor_cache.clear()
        _anchor_cache[key] = t
    return t


def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The float64 taps scipy.ndimage.gaussian_filter builds (``_gaussian_kernel1d``), which is what
    skimage.filters.gaussian(sigma, preserve_range=True) of utils/grasp_eval.py:198 runs with."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return np.ascontiguousarray(w / w.sum(), dtype=np.float64)


def gaussian_batched(maps: torch.Tensor, sigma: float = 2.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[P, H, W] float32 CUDA maps -> Gaussian-smoothed maps with scipy/skimage semantics (edge-replicate, separable,
    float64 accumulation, float32 result per pass)."""
    lib = L.lib()
    assert maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 3 and maps.is_contiguous()
    taps = gaussian_taps(sigma)
    P, H, W = maps.shape
    out = torch.empty_like(maps) if out is None else out
    radius = (len(taps) - 1) // 2
    # out of place with sigma = 2 the library runs both passes in one kernel; otherwise it needs the intermediate buffer
    two_pass = (out.data_ptr() == maps.data_ptr() or radius != 8 or "CROG_GAUSSIAN_TWOPASS" in os.environ
                or "CROG_GAUSSIAN_GENERIC" in os.environ)
    tmp = torch.empty_like(maps) if two_pass else None
    with torch.cuda.device(maps.device):
        L.check(lib.crog_gaussian(maps.data_ptr(), tmp.data_ptr() if two_pass else None, out.data_ptr(), P, H, W, taps.ctypes.data,
                                  radius, None, 1, 0, L.stream_ptr()))
    return out


def _ssg_detect(cfg, cls: torch.Tensor, box: torch.Tensor, anchors: torch.Tensor, score_thr2: float = 0.3):
    """grasp_eval.py:113-150 on the device for one image: returns (boxes [N,4], det_n, det_anchor, det_class, det_score)."""
    lib = L.lib()
    N, nc = cls.shape
    dev = cls.device
    md = int(cfg.max_detections)
    keep = torch.empty((N,), dtype=torch.int32, device=dev)
    boxes = torch.empty((N, 4), dtype=torch.float32, device=dev)
    det_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    det_anchor = torch.empty((md,), dtype=torch.int32, device=dev)
    det_class = torch.empty((md,), dtype=torch.int32, device=dev)
    det_score = torch.empty((md,), dtype=torch.float32, device=dev)
    ws = torch.empty((int(L.load().crog_ssg_nms_workspace_bytes(nc, int(cfg.top_k))),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_ssg_detect(cls.data_ptr(), box.data_ptr(), anchors.data_ptr(), N, nc, float(cfg.nms_score_thre),
                                    float(cfg.nms_iou_thre), int(cfg.top_k), md, float(score_thr2), keep.data_ptr(), boxes.data_ptr(),
                                    det_n.data_ptr(), det_anchor.data_ptr(), det_class.data_ptr(), det_score.data_ptr(),
                                    ws.data_ptr(), L.stream_ptr()))
    return boxes, det_n, det_anchor, det_class, det_score, keep


def fast_nms(cfg, box_pred_kept, cls_pred_kept, ins_coef_pred_kept, grasp_coef_pred_kept):
    """utils/grasp_eval.py:55-93 with the reference's signature: ``box_pred_kept`` [n,4] decoded boxes, ``cls_pred_kept``
    [num_fg_classes, n] scores -> (class_ids, scores, boxes, coefs, grasp_coefs) of the surviving detections, best first.
    Ties in either sort go to the lower index (the reference leaves them to torch.sort)."""
    lib = L.lib()
    dev = _dev()
    box = torch.as_tensor(box_pred_kept).to(dev, torch.float32).contiguous()
    sc = torch.as_tensor(cls_pred_kept).to(dev, torch.float32)
    coef, gco = torch.as_tensor(ins_coef_pred_kept).to(dev), torch.as_tensor(grasp_coef_pred_kept).to(dev)
    ncf, n = sc.shape
    md = int(cfg.max_detections)
    cls_full = torch.cat([torch.zeros((1, n), device=dev), sc]).t().contiguous()  # [n, 1 + fg classes]; column 0 is ignored
    keep = torch.ones((max(n, 1),), dtype=torch.int32, device=dev)
    det_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    det_anchor = torch.empty((md,), dtype=torch.int32, device=dev)
    det_class = torch.empty((md,), dtype=torch.int32, device=dev)
    det_score = torch.empty((md,), dtype=torch.float32, device=dev)
    ws = torch.empty((int(L.load().crog_ssg_nms_workspace_bytes(ncf + 1, int(cfg.top_k))),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_ssg_fast_nms(cls_full.data_ptr(), keep.data_ptr(), box.data_ptr(), n, ncf + 1, float(cfg.nms_iou_thre),
                                      int(cfg.top_k), md, float("inf"), det_n.data_ptr(), det_anchor.data_ptr(), det_class.data_ptr(),
                                      det_score.data_ptr(), ws.data_ptr(), L.stream_ptr()))
    k = int(det_n.item())
    idx = det_anchor[:k].long()
    return det_class[:k].long(), det_score[:k], box[idx], coef[idx], gco[idx]


def ssg_masks_device(cfg, output_b, boxes, det_anchor, n: int, ori_size):
    """grasp_eval.py:171-198 for one image with ``n`` detections, on the device: returns ``hr`` [5, n, ori_h, ori_w] float32
    (instance mask (0/1), quality — already Gaussian-smoothed —, sin, cos, width)."""
    lib = L.lib()
    protos, coef, gco = output_b
    ori_h, ori_w = int(ori_size[0]), int(ori_size[1])
    S = max(ori_h, ori_w)
    h, w, npz = protos.shape
    dev = protos.device
    lowres = torch.empty((n, 5, h, w), dtype=torch.float32, device=dev)
    hr = torch.empty((5, n, ori_h, ori_w), dtype=torch.float32, device=dev)
    if n == 0:
        return hr
    qraw = torch.empty((n, ori_h, ori_w), dtype=torch.float32, device=dev)  # un-smoothed quality maps
    n_dev = torch.full((1,), n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.crog_ssg_masks(protos.data_ptr(), h, w, npz, coef.data_ptr(), gco.data_ptr(), boxes.data_ptr(),
                                   det_anchor.data_ptr(), n_dev.data_ptr(), n, lowres.data_ptr(), hr.data_ptr(), qraw.data_ptr(), n,
                                   ori_h, ori_w, S, L.stream_ptr()))
    gaussian_batched(qraw, 2.0, out=hr[1])
    return hr


@torch.no_grad()
def ssg_post_processing(cfg, output_dict, data_dict):
    """utils/grasp_eval.py:100-221 (batch size 1, like the reference): score filter, box decode, Fast NMS, prototype
    assembly + crop, bilinear resize to the original image, Gaussian smoothing of the quality maps and grasp decode —
    all on the B200; the returned dict has the reference's keys and host types."""
    dev = _dev()
    ori_h, ori_w = data_dict["ori_size"]
    f = lambda k: torch.as_tensor(output_dict[k]).to(dev, torch.float32).squeeze(0).contiguous()
    protos, cls, box, coef, gco = f("protos"), f("cls_pred"), f("box_pred"), f("ins_coef_pred"), f("grasp_coef_pred")
    if cls.dim() != 2:
        raise L.CrogError("ssg_post_processing handles one image per call (batch size 1), like the reference")
    anchors = _anchors_dev(output_dict["anchors"], dev)
    boxes, det_n, det_anchor, det_class, det_score, _ = _ssg_detect(cfg, cls, box, anchors)
    n = int(det_n.item())  # dynamic output shapes, as in the reference
    hr = ssg_masks_device(cfg, (protos, coef, gco), boxes, det_anchor, n, (ori_h, ori_w))
    ins, qua, sin, cos, wid = hr[0], hr[1], hr[2], hr[3], hr[4]
    top1, top5 = [], []
    ang = torch.empty_like(sin)
    if n > 0:
        ang = angle_map(sin, cos)
        _, npk, grasps = detect_grasps_batched(qua, sin, cos, wid, 5)
        npk_h, g_h = npk.cpu().tolist(), grasps.cpu().tolist()
        for i in range(n):
            rows = [[r[0], r[1], r[2], 20, r[4]] for r in g_h[i][:npk_h[i]]]
            top5.append(rows)
            top1.append(rows[:1])  # the greedy peak order does not depend on K: top-1 is the first of top-5
    idx = det_anchor[:n].long()
    return {
        "cls": (det_class[:n].long() + 1).cpu().numpy(),
        "bboxes": boxes[idx].cpu().numpy() * np.array([ori_w, ori_w, ori_w, ori_w]),
        "ins_masks": ins.cpu().numpy(),
        "grasps_top1": top1,
        "grasps_top5": top5,
        "grasp_masks": (qua.cpu().numpy(), ang.cpu().numpy(), wid.cpu().numpy()),
    }


@torch.no_grad()
def ssg_post_processing_batched(cfg, output_dict, ori_size=(480, 640)):
    """Device-resident batched form used by the engine / benchmark: the per-image stages of ssg_post_processing for every
    sample of a batch with ONE host synchronisation (the detection counts) and nothing else leaving HBM.  Every stage is a
    batch-wide launch: box decode + Fast NMS (three kernels, grid dimension = image); after the counts are known, the masks
    of all instances into ONE map-major tensor ``[5, sum(n), H, W]`` (two kernels), one Gaussian, one peak decode.
    Returns a list of per-sample dicts of CUDA tensors (views): n, cls, boxes, scores, hr [5,n,H,W], n_peaks [n],
    grasps [n,5,5]."""
    lib = L.lib()
    protos, cls, box = output_dict["protos"], output_dict["cls_pred"], output_dict["box_pred"]
    coef, gco = output_dict["ins_coef_pred"], output_dict["grasp_coef_pred"]
    dev = protos.device
    f32c = lambda t: t.to(dev, torch.float32).contiguous()
    protos, cls, box, coef, gco = f32c(protos), f32c(cls), f32c(box), f32c(coef), f32c(gco)
    anchors = _anchors_dev(output_dict["anchors"], dev)
    B, N, nc = cls.shape
    h, w, npz = protos.shape[1:]
    md = int(cfg.max_detections)
    ori_h, ori_w = int(ori_size[0]), int(ori_size[1])
    S = max(ori_h, ori_w)
    ws_bytes = ((int(L.load().crog_ssg_nms_workspace_bytes(nc, int(cfg.top_k))) + 255) // 256) * 256
    keep = torch.empty((B, N), dtype=torch.int32, device=dev)
    boxes = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    det_n = torch.zeros((B,), dtype=torch.int32, device=dev)
    det_anchor = torch.empty((B, md), dtype=torch.int32, device=dev)
    det_class = torch.empty((B, md), dtype=torch.int32, device=dev)
    det_score = torch.empty((B, md), dtype=torch.float32, device=dev)
    ws = torch.empty((B, ws_bytes), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        s = L.stream_ptr()
        # box decode + Fast NMS of every image in three launches (grid dimension = image)
        L.check(lib.crog_ssg_detect_batched(cls.data_ptr(), box.data_ptr(), anchors.data_ptr(), B, N, nc, float(cfg.nms_score_thre),
                                            float(cfg.nms_iou_thre), int(cfg.top_k), md, 0.3, keep.data_ptr(), boxes.data_ptr(),
                                            det_n.data_ptr(), det_anchor.data_ptr(), det_class.data_ptr(), det_score.data_ptr(),
                                            ws.data_ptr(), ws_bytes, s))
        counts = det_n.cpu().tolist()  # the one sync: output shapes are data dependent, as in the reference
        offs = [0]
        for n in counts:
            offs.append(offs[-1] + n)
        tot = offs[-1]
        hr = torch.empty((5, max(tot, 1), ori_h, ori_w), dtype=torch.float32, device=dev)
        lowres = torch.empty((max(tot, 1), 5, h, w), dtype=torch.float32, device=dev)
        qraw = torch.empty((max(tot, 1), ori_h, ori_w), dtype=torch.float32, device=dev)  # un-smoothed quality maps
        if tot > 0:
            # all instances of the batch in two launches: instance i = detection inst[1, i] of image inst[0, i]
            inst = torch.from_numpy(np.stack([np.repeat(np.arange(B, dtype=np.int32), counts),
                                              np.concatenate([np.arange(n, dtype=np.int32) for n in counts])])).to(dev)
            L.check(lib.crog_ssg_masks_batched(protos.data_ptr(), h, w, npz, coef.data_ptr(), gco.data_ptr(), boxes.data_ptr(),
                                               det_anchor.data_ptr(), N, md, inst[0].data_ptr(), inst[1].data_ptr(), tot,
                                               lowres.data_ptr(), hr.data_ptr(), qraw.data_ptr(), ori_h, ori_w, S, s))
        if tot > 0:
            gaussian_batched(qraw[:tot], 2.0, out=hr[1, :tot])  # smoothed out of place into the quality plane
            _, npk, grasps = detect_grasps_batched(hr[1, :tot], hr[2, :tot], hr[3, :tot], hr[4, :tot], 5)
        else:
            npk = torch.zeros((0,), dtype=torch.int32, device=dev)
            grasps = torch.zeros((0, 5, 5), dtype=torch.float64, device=dev)
        # per-sample views of batch-wide results (one gather for all boxes instead of one per image)
        sel = det_anchor.long().clamp_(0, N - 1)  # entries past an image's count are unused
        boxes_sel = torch.gather(boxes, 1, sel[..., None].expand(-1, -1, 4))
        cls1 = det_class + 1
    out = []
    for b in range(B):
        n, o = counts[b], offs[b]
        out.append({"n": n, "cls": cls1[b, :n], "boxes": boxes_sel[b, :n], "scores": det_score[b, :n],
                    "hr": hr[:, o:o + n], "n_peaks": npk[o:o + n], "grasps": grasps[o:o + n]})
    return out
