"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference SSG
inference path — ``SSG.forward`` (model/ssg.py:248-279) and ``ssg_post_processing`` / ``fast_nms``
(utils/grasp_eval.py:55-221) with ``utils/box_utils.py`` helpers.

Pinning: ``oracle/make_golden_ssg.py`` imports the unmodified reference ``model.ssg.SSG`` from
/root/reference, loads the same synthetic state-dict with ``strict=True`` and asserts this forward agrees
to fp32 round-off before writing tests/golden/ssg_*.npz.  The reference's post-processing cannot be
imported here (utils/grasp_eval.py needs scikit-image / matplotlib); it is restated line
by line; ``gaussian`` follows scikit-image 0.20.0 -> ``scipy.ndimage.gaussian_filter(mode='nearest',
truncate=4.0)`` and is pinned against the installed scipy in tests/test_oracle_ssg.py.  The torch-only
parts (fast_nms, box_iou, crop, sanitize_coordinates) ARE importable from the reference and are compared
with it by make_golden_ssg.py.  PARITY UNPINNED for the scikit-image call sites (peak_local_max), as for
the CROG tail.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import grasp_tail as T

BN_EPS = 1e-5


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _bottleneck(sd, p, x, stride):
    """model/ssg.py:15-50 (stride lives in the 3x3 conv)."""
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + res)


def backbone(sd, cfg, img):
    """model/ssg.py:97-110."""
    x = F.relu(_bn(sd, "backbone.bn1", F.conv2d(img, sd["backbone.conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nb in enumerate(cfg.resnet_layers):
        for bi in range(nb):
            x = _bottleneck(sd, f"backbone.layers.{li}.{bi}", x, 2 if (li > 0 and bi == 0) else 1)
        outs.append(x)
    return outs


def _cb(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def fpn(sd, c3, c4, c5):
    """model/ssg.py:189-205."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    p5_1 = _cb(sd, "fpn.lat_layers.2", c5)
    p4_1 = _cb(sd, "fpn.lat_layers.1", c4) + up(p5_1)
    p3_1 = _cb(sd, "fpn.lat_layers.0", c3) + up(p4_1)
    p5 = F.relu(_cb(sd, "fpn.pred_layers.2.0", p5_1, padding=1))
    p4 = F.relu(_cb(sd, "fpn.pred_layers.1.0", p4_1, padding=1))
    p3 = F.relu(_cb(sd, "fpn.pred_layers.0.0", p3_1, padding=1))
    p6 = F.relu(_cb(sd, "fpn.downsample_layers.0.0", p5, stride=2, padding=1))
    p7 = F.relu(_cb(sd, "fpn.downsample_layers.1.0", p6, stride=2, padding=1))
    return [p3, p4, p5, p6, p7]


def proto_net(sd, x):
    """model/ssg.py:150-169."""
    for i in (0, 2, 4):
        x = F.relu(_cb(sd, f"proto_net.proto1.{i}", x, padding=1))
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.relu(_cb(sd, "proto_net.proto2.0", x, padding=1))
    return F.relu(_cb(sd, "proto_net.proto2.2", x))


def prediction(sd, cfg, x):
    """model/ssg.py:136-147."""
    B = x.shape[0]
    x = F.relu(_cb(sd, "prediction_layers.upfeature.0", x, padding=1))
    conf = _cb(sd, "prediction_layers.conf_layer", x, padding=1).permute(0, 2, 3, 1).reshape(B, -1, cfg.num_classes)
    box = _cb(sd, "prediction_layers.bbox_layer", x, padding=1).permute(0, 2, 3, 1).reshape(B, -1, 4)
    coef = torch.tanh(_cb(sd, "prediction_layers.coef_layer.0", x, padding=1)).permute(0, 2, 3, 1).reshape(B, -1, cfg.num_protos)
    gco = torch.tanh(_cb(sd, "prediction_layers.grasp_coef_layer.0", x, padding=1)).permute(0, 2, 3, 1).reshape(B, -1, 4, cfg.num_protos)
    return conf, box, coef, gco


@torch.no_grad()
def ssg_forward(sd: Dict[str, torch.Tensor], cfg, rgb: torch.Tensor, depth: torch.Tensor, keep: bool = False):
    """model/ssg.py:248-279 (eval branch)."""
    from crog_b200.model.ssg_anchors import make_all_anchors

    img = torch.cat([rgb, depth], 1) if cfg.with_depth else rgb
    feats = backbone(sd, cfg, img)
    pyr = fpn(sd, *feats[1:4])
    protos = proto_net(sd, pyr[0]).permute(0, 2, 3, 1).contiguous()
    parts = [prediction(sd, cfg, p) for p in pyr]
    cls, box, coef, gco = [torch.cat([p[i] for p in parts], 1) for i in range(4)]
    out = {"anchors": make_all_anchors(cfg), "protos": protos, "cls_pred": F.softmax(cls, -1), "box_pred": box,
           "ins_coef_pred": coef, "grasp_coef_pred": gco}
    if keep:
        return out, {"c2": feats[0], "c3": feats[1], "c4": feats[2], "c5": feats[3], "p3": pyr[0], "p5": pyr[2], "p7": pyr[4]}
    return out


# ====================================================================== post-processing
def box_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils/box_utils.py:8-37, batched [n, A, 4] x [n, B, 4] point-form boxes."""
    max_xy = torch.min(a[:, :, None, 2:], b[:, None, :, 2:])
    min_xy = torch.max(a[:, :, None, :2], b[:, None, :, :2])
    inter = torch.clamp(max_xy - min_xy, min=0)
    ia = inter[..., 0] * inter[..., 1]
    aa = ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]))[:, :, None]
    ab = ((b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]))[:, None, :]
    return ia / (aa + ab - ia)


def fast_nms(cfg, box, cls, coef, gco):
    """utils/grasp_eval.py:55-93.  Sorting ties are broken by the lower index (stable), which is what
    torch's CPU sort does; the reference leaves it unspecified."""
    cls, idx = cls.sort(dim=1, descending=True, stable=True)
    idx = idx[:, :cfg.top_k]
    cls = cls[:, :cfg.top_k]
    nc, nd = idx.shape
    box = box[idx.reshape(-1)].reshape(nc, nd, 4)
    coef = coef[idx.reshape(-1)].reshape(nc, nd, -1)
    gco = gco[idx.reshape(-1)].reshape(nc, nd, 4, -1)
    iou = box_iou(box, box)
    iou.triu_(diagonal=1)
    iou_max, _ = iou.max(dim=1)
    keep = iou_max <= cfg.nms_iou_thre
    class_ids = torch.arange(nc)[:, None].expand_as(keep)[keep]
    cls, box, coef, gco = cls[keep], box[keep], coef[keep], gco[keep]
    cls, idx = cls.sort(dim=0, descending=True, stable=True)
    idx = idx[:cfg.max_detections]
    return class_ids[idx], cls[:cfg.max_detections], box[idx], coef[idx], gco[idx]


def decode_boxes(anchors: torch.Tensor, box: torch.Tensor) -> torch.Tensor:
    """utils/grasp_eval.py:133-137."""
    b = torch.cat((anchors[:, :2] + box[:, :2] * 0.1 * anchors[:, 2:], anchors[:, 2:] * torch.exp(box[:, 2:] * 0.2)), 1)
    b[:, :2] -= b[:, 2:] / 2
    b[:, 2:] += b[:, :2]
    return torch.clip(b, min=0., max=1.)


def crop(masks: torch.Tensor, boxes: torch.Tensor, padding: int = 1) -> torch.Tensor:
    """utils/box_utils.py:150-171 (+ sanitize_coordinates :120-135); masks [h, w, n]."""
    h, w, n = masks.shape

    def san(a, b, size):
        a, b = a * size, b * size
        lo, hi = torch.min(a, b), torch.max(a, b)
        return torch.clamp(lo - padding, min=0), torch.clamp(hi + padding, max=size)

    x1, x2 = san(boxes[:, 0], boxes[:, 2], w)
    y1, y2 = san(boxes[:, 1], boxes[:, 3], h)
    rows = torch.arange(w, dtype=x1.dtype).view(1, -1, 1).expand(h, w, n)
    cols = torch.arange(h, dtype=x1.dtype).view(-1, 1, 1).expand(h, w, n)
    m = (rows >= x1.view(1, 1, -1)) * (rows < x2.view(1, 1, -1)) * (cols >= y1.view(1, 1, -1)) * (cols < y2.view(1, 1, -1))
    return masks * m.float()


def gaussian_f32(img: np.ndarray, sigma: float = 2.0, truncate: float = 4.0) -> np.ndarray:
    """skimage.filters.gaussian(img, sigma, preserve_range=True) on a 2-D float32 map =
    scipy.ndimage.gaussian_filter(mode='nearest', truncate=4): radius int(truncate*sigma+0.5), float64 weights
    exp(-x^2/2s^2)/sum, axis 0 then axis 1, each pass accumulated in float64 in scipy's symmetric order
    (centre tap, then pairs from the outermost inwards) and rounded to float32 once."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    w = w / w.sum()
    out = np.asarray(img, np.float32)
    for axis in (0, 1):
        a = np.moveaxis(out, axis, 0).astype(np.float64)
        n = a.shape[0]
        idx = np.clip(np.arange(-r, n + r), 0, n - 1)
        p = a[idx]  # edge-replicated
        acc = p[r:r + n] * w[r]
        for j in range(-r, 0):
            acc = acc + (p[r + j:r + j + n] + p[r - j:r - j + n]) * w[j + r]
        out = np.moveaxis(acc.astype(np.float32), 0, axis)
    return np.ascontiguousarray(out)


@torch.no_grad()
def ssg_post_processing(cfg, output_dict, data_dict, keep: bool = False):
    """utils/grasp_eval.py:100-221, batch size 1."""
    ori_h, ori_w = data_dict["ori_size"]
    input_size = max(ori_h, ori_w)
    protos = output_dict["protos"].squeeze().float()
    cls_pred = output_dict["cls_pred"].squeeze().float()
    box_pred = output_dict["box_pred"].squeeze().float()
    coef = output_dict["ins_coef_pred"].squeeze().float()
    gco = output_dict["grasp_coef_pred"].squeeze().float()
    anchors = torch.tensor(output_dict["anchors"], dtype=torch.float32).reshape(-1, 4)
    cls_pred = cls_pred.transpose(1, 0).contiguous()[1:, :]
    cls_max, _ = torch.max(cls_pred, dim=0)
    keep0 = cls_max > cfg.nms_score_thre
    boxes = decode_boxes(anchors[keep0], box_pred[keep0])
    class_ids, scores, boxes, coef_k, gco_k = fast_nms(cfg, boxes, cls_pred[:, keep0], coef[keep0], gco[keep0])
    k2 = scores > 0.3
    if k2.any():
        class_ids, scores, boxes, coef_k, gco_k = class_ids[k2], scores[k2], boxes[k2], coef_k[k2], gco_k[k2]
    class_ids = (class_ids + 1).numpy()
    lr = [torch.sigmoid(protos @ coef_k.t()), torch.sigmoid(protos @ gco_k[:, 0].t()), protos @ gco_k[:, 1].t(),
          protos @ gco_k[:, 2].t(), torch.sigmoid(protos @ gco_k[:, 3].t())]
    lr = [crop(m.contiguous(), boxes).permute(2, 0, 1) for m in lr]
    n = lr[0].shape[0]
    if n > 0:
        hr = [F.interpolate(m.unsqueeze(0), (input_size, input_size), mode="bilinear", align_corners=False).squeeze(0) for m in lr]
    else:
        hr = [torch.zeros((0, input_size, input_size)) for _ in lr]
    hr[0] = (hr[0] > 0.5).float()
    ins, qua, sin, cos, wid = [m[:, :ori_h, :ori_w].contiguous().numpy() for m in hr]
    qua_raw = qua.copy()
    ang = []
    for i in range(n):
        qua[i] = gaussian_f32(qua[i], 2.0)
        ang.append((np.arctan2(sin[i].astype(np.float64), cos[i].astype(np.float64)).astype(np.float32) * np.float32(0.5)))
    ang = np.asarray(ang)
    boxes_px = boxes.numpy() * np.array([ori_w, ori_w, ori_w, ori_w])
    top1, top5 = [], []
    for i in range(n):
        top1.append(T.detect_grasps(qua[i], sin[i], cos[i], wid[i], 1)[0])
        top5.append(T.detect_grasps(qua[i], sin[i], cos[i], wid[i], 5)[0])
    out = {"cls": class_ids, "bboxes": boxes_px, "ins_masks": ins, "grasps_top1": top1, "grasps_top5": top5,
           "grasp_masks": (qua, ang, wid)}
    if keep:
        out["_debug"] = {"scores": scores.numpy(), "boxes_rel": boxes.numpy(), "qua_raw": qua_raw, "sin": sin, "cos": cos,
                         "lowres": [m.numpy() for m in lr]}
    return out
