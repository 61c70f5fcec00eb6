"""Seeded synthetic configs, weights, inputs and ground truth.

Everything here is generated on the CPU with explicit ``torch.Generator`` /
``numpy.random.Generator`` seeds so the same bytes are produced in the build
container (where the reference is imported to make golden vectors) and on the GPU
box (where ``/root/reference`` does not exist).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Tuple

import numpy as np
import torch

from .spec import crog_tensor_specs

SOT_TOKEN = 49406
EOT_TOKEN = 49407


def default_cfg(word_len: int = 17, **over) -> SimpleNamespace:
    """Model keys of config/OCID-VLG/crog_multiple_r50.yaml:8-22,46-48."""
    cfg = SimpleNamespace(
        clip_pretrain="synthetic", input_size=416, word_len=word_len, word_dim=1024, vis_dim=512,
        fpn_in=[512, 1024, 1024], fpn_out=[256, 512, 1024], num_layers=3, num_head=8, dim_ffn=2048,
        dropout=0.1, intermediate=False, use_contrastive=True, use_pretrained_clip=False,
        use_grasp_masks=True, lr_multi=0.1, base_lr=1e-4)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def make_state_dict(cfg, seed: int = 0, mode: str = "perturbed") -> Dict[str, torch.Tensor]:
    """Deterministic random weights for every tensor of the CROG state-dict.

    mode="init":      distributions of the reference's random init
                      (model/clip.py:378-420 + torch defaults): BN = identity, every
                      ``bn3.weight`` of the residual stages is zero.
    mode="perturbed": He-scaled convs and randomised BN affine/statistics so that all
                      16 bottleneck branches and both BN terms are exercised
                      (SURVEY.md §7 "Random-init hides bugs").
    """
    assert mode in ("init", "perturbed")
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def randn(shape, std):
        return torch.randn(shape, generator=g) * std

    def rand(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    for s in crog_tensor_specs(cfg):
        n = s.name
        if s.role == "conv" or s.role == "linear_w" or s.role == "proj":
            if mode == "init":
                if ".attnpool." in n and n.endswith("_proj.weight"):
                    t = randn(s.shape, 2048 ** -0.5)
                elif "backbone.transformer" in n:
                    wdt = 512
                    std = {"in_proj_weight": wdt ** -0.5, "out_proj.weight": wdt ** -0.5 * 24 ** -0.5,
                           "c_fc.weight": (2 * wdt) ** -0.5, "c_proj.weight": wdt ** -0.5 * 24 ** -0.5}
                    t = randn(s.shape, next(v for k, v in std.items() if n.endswith(k)))
                elif s.role == "proj":
                    t = randn(s.shape, 512 ** -0.5)
                else:
                    b = 1.0 / math.sqrt(s.fan_in)
                    t = rand(s.shape, -b, b)
            else:
                gain = 2.0 if s.role == "conv" else 1.0
                if n == "proj.txt.weight":
                    gain = 1.0 / 256.0  # keeps the dynamic-conv logits O(1..10)
                t = randn(s.shape, math.sqrt(gain / s.fan_in))
        elif s.role == "bias":
            b = 1.0 / math.sqrt(max(s.fan_in, 1))
            t = rand(s.shape, -b, b) if mode == "init" else randn(s.shape, 0.05)
        elif s.role == "bn_w":
            if mode == "init":
                zero = "backbone.visual.layer" in n and n.endswith("bn3.weight")
                t = torch.zeros(s.shape) if zero else torch.ones(s.shape)
            elif "backbone.visual.layer" in n and n.endswith("bn3.weight"):
                t = rand(s.shape, 0.1, 0.5)  # non-zero residual branches without blowing up the trunk
            else:
                t = rand(s.shape, 0.5, 1.5)
        elif s.role == "bn_b":
            t = torch.zeros(s.shape) if mode == "init" else randn(s.shape, 0.1)
        elif s.role == "bn_mean":
            t = torch.zeros(s.shape) if mode == "init" else randn(s.shape, 0.1)
        elif s.role == "bn_var":
            t = torch.ones(s.shape) if mode == "init" else rand(s.shape, 0.5, 1.5)
        elif s.role == "bn_count":
            t = torch.zeros((), dtype=torch.int64)
        elif s.role == "ln_w":
            t = torch.ones(s.shape) if mode == "init" else rand(s.shape, 0.8, 1.2)
        elif s.role == "ln_b":
            t = torch.zeros(s.shape) if mode == "init" else randn(s.shape, 0.05)
        elif s.role == "embed":
            t = randn(s.shape, 0.02 if mode == "init" else 0.5)
        elif s.role == "pos":
            if "attnpool" in n:
                t = randn(s.shape, s.shape[1] ** -0.5 if mode == "init" else 0.3)
            else:
                t = randn(s.shape, 0.01 if mode == "init" else 0.3)
        elif s.role == "scalar":
            t = torch.tensor(math.log(1 / 0.07))
        else:  # pragma: no cover
            raise KeyError(s.role)
        sd[n] = t.contiguous()
    return sd


def make_inputs(batch: int, word_len: int, size: int = 416, seed_img: int = 1,
                seed_txt: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Config-1/2 inputs: ``img`` ~ N(0,1) fp32 B×3×S×S; ``word`` int64 B×L =
    ``[SOT, t_1..t_n, EOT, 0...]`` with n in [3, 12] (EOT is the arg-max, pads exercise
    ``pad_mask``; model/crog.py:55, model/clip.py:451)."""
    gi = torch.Generator().manual_seed(seed_img)
    img = torch.randn((batch, 3, size, size), generator=gi)
    gt = torch.Generator().manual_seed(seed_txt)
    word = torch.zeros((batch, word_len), dtype=torch.int64)
    for b in range(batch):
        n = int(torch.randint(3, 13, (1,), generator=gt))
        n = min(n, word_len - 2)
        toks = torch.randint(1, SOT_TOKEN, (n,), generator=gt)
        word[b, 0] = SOT_TOKEN
        word[b, 1:1 + n] = toks
        word[b, 1 + n] = EOT_TOKEN
    return img, word


def make_gt_rects(batch: int, max_gt: int = 64, seed: int = 4, size: int = 416
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Config-3 ground truth: per sample M~U{1..max_gt} rectangles
    ``[cx, cy, w, h, theta_deg, cls]`` float64, padded to ``max_gt`` rows.
    w>100 and h!=20 exercise the in-place clip/overwrite of grasp_eval.py:367-368."""
    rng = np.random.default_rng(seed)
    gt = np.zeros((batch, max_gt, 6), dtype=np.float64)
    cnt = rng.integers(1, max_gt + 1, size=batch).astype(np.int32)
    lo, hi = 40.0, size - 40.0
    for b in range(batch):
        m = int(cnt[b])
        gt[b, :m, 0] = rng.uniform(lo, hi, m)
        gt[b, :m, 1] = rng.uniform(lo, hi, m)
        gt[b, :m, 2] = rng.uniform(10, 120, m)
        gt[b, :m, 3] = rng.uniform(10, 40, m)
        gt[b, :m, 4] = rng.uniform(-90, 90, m)
        gt[b, :m, 5] = 1.0
    return gt, cnt


def make_tail_maps(n: int, kind: str = "blobs", seed: int = 7, size: int = 416
                   ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Config-5 synthetic quality / sin / cos / width maps, float32 n×S×S.

    kind="blobs":  q = clip(sum of <=8 Gaussians + N(0,0.01), 0, 1); smooth angle field.
    kind="stress": q ~ U(0,1) iid (thousands of candidate peaks, near-ties); every 20th
                   map is quantised to 1/16 to create plateaus (tie / spacing rules).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    q = np.empty((n, size, size), np.float32)
    s = np.empty_like(q); c = np.empty_like(q); w = np.empty_like(q)
    for i in range(n):
        if kind == "blobs":
            acc = np.zeros((size, size), np.float32)
            for _ in range(int(rng.integers(1, 9))):
                a = rng.uniform(0.45, 1.0); sg = rng.uniform(3, 10)
                mx, my = rng.uniform(10, size - 10, 2)
                acc += (a * np.exp(-((xx - mx) ** 2 + (yy - my) ** 2) / (2 * sg * sg))).astype(np.float32)
            acc += rng.normal(0, 0.01, acc.shape).astype(np.float32)
            q[i] = np.clip(acc, 0, 1)
        elif kind == "stress":
            m = rng.random((size, size), dtype=np.float32)
            if i % 20 == 19:
                m = np.floor(m * 16).astype(np.float32) / 16
            q[i] = m
        else:
            raise ValueError(kind)
        kx, ky, ph = rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-np.pi, np.pi)
        phi = (kx * xx + ky * yy + ph).astype(np.float32)
        s[i] = np.sin(2 * phi) + rng.normal(0, 0.05, phi.shape).astype(np.float32)
        c[i] = np.cos(2 * phi) + rng.normal(0, 0.05, phi.shape).astype(np.float32)
        w[i] = (0.5 + 0.5 * np.sin(rng.uniform(0.005, 0.03) * xx + rng.uniform(0.005, 0.03) * yy)).astype(np.float32)
    return q, s, c, w


# ====================================================================== SSG (config 4)
def ssg_cfg(**over) -> SimpleNamespace:
    """Model / post-processing keys of config/OCID-Grasp/ssg_r50.yaml:6-31,52-57."""
    cfg = SimpleNamespace(
        backbone="resnet", resnet_layers=[3, 4, 6, 3], path_to_pretrained_resnet=None, resume=None, with_depth=True,
        fpn_in_channels=[512, 1024, 2048], num_protos=32, num_classes=32, aspect_ratios=[1, 0.5, 2],
        anchor_strides=[8, 16, 32, 64, 128], with_grasp_masks=True, img_size=544,
        nms_score_thre=0.05, nms_iou_thre=0.5, top_k=200, max_detections=100)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def make_ssg_state_dict(cfg, seed: int = 0, mode: str = "perturbed") -> Dict[str, torch.Tensor]:
    """mode="init": the reference's xavier-uniform convs, zero biases, identity BN (ssg.py:241-245).
    mode="perturbed": He-scaled convs, random biases and BN affine / statistics (every term exercised)."""
    from .spec import ssg_tensor_specs

    assert mode in ("init", "perturbed")
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for s in ssg_tensor_specs(cfg):
        if s.role == "conv":
            if mode == "init":
                co, ci, kh, kw = s.shape
                bound = math.sqrt(6.0 / (ci * kh * kw + co * kh * kw))
                t = (torch.rand(s.shape, generator=g) * 2 - 1) * bound
            else:
                gain = 1.0 if ("prediction_layers" in s.name and "upfeature" not in s.name) or "proto2.2" in s.name else 2.0
                t = torch.randn(s.shape, generator=g) * math.sqrt(gain / s.fan_in)
        elif s.role == "bias":
            t = torch.zeros(s.shape) if mode == "init" else torch.randn(s.shape, generator=g) * 0.05
        elif s.role == "bn_w":
            if mode == "init":
                t = torch.ones(s.shape)
            elif s.name.endswith("bn3.weight"):
                t = torch.rand(s.shape, generator=g) * 0.4 + 0.1
            else:
                t = torch.rand(s.shape, generator=g) + 0.5
        elif s.role in ("bn_b", "bn_mean"):
            t = torch.zeros(s.shape) if mode == "init" else torch.randn(s.shape, generator=g) * 0.1
        elif s.role == "bn_var":
            t = torch.ones(s.shape) if mode == "init" else torch.rand(s.shape, generator=g) + 0.5
        elif s.role == "bn_count":
            t = torch.zeros((), dtype=torch.int64)
        else:  # pragma: no cover
            raise KeyError(s.role)
        sd[s.name] = t.contiguous()
    return sd


def make_ssg_inputs(batch: int, size: int = 544, seed: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Config-4 inputs: rgb ~ U(0,1) B x 3 x S x S, depth ~ U(0,1) B x 1 x S x S."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((batch, 3, size, size), generator=g), torch.rand((batch, 1, size, size), generator=g)


def make_ssg_output_dict(cfg, n_confident: int = 8, seed: int = 6, proto_hw: int = 136) -> Dict[str, torch.Tensor]:
    """A synthetic single-image ``output_dict`` with ``n_confident`` confident, well separated detections (random-init
    scores never pass 0.3, so post-processing is exercised on injected detections)."""
    from .model.ssg_anchors import make_all_anchors

    g = torch.Generator().manual_seed(seed)
    anchors = torch.tensor(make_all_anchors(cfg), dtype=torch.float32).view(-1, 4)
    N, nc, npz = anchors.shape[0], cfg.num_classes, cfg.num_protos
    logits = torch.randn((N, nc), generator=g) * 0.5
    logits[:, 0] += 6.0  # background dominates
    box = torch.randn((N, 4), generator=g) * 0.3
    coef = torch.tanh(torch.randn((N, npz), generator=g))
    gcoef = torch.tanh(torch.randn((N, 4, npz), generator=g))
    # confident anchors on the stride-16/32 levels, spread over the image, plus near-duplicates for NMS to remove
    lvl0 = 3 * (proto_hw // 2) ** 2
    picks = lvl0 + torch.randperm(3 * (proto_hw // 4) ** 2, generator=g)[:n_confident]
    for j, a in enumerate(picks.tolist()):
        cls = 1 + (j * 7) % (nc - 1)
        logits[a, cls] += 9.0 + 0.1 * j
        if a + 3 < N:  # neighbouring cell, same ratio: overlaps heavily -> suppressed by Fast NMS
            logits[a + 3, cls] += 8.0
            box[a + 3] = box[a]
    protos = torch.relu(torch.randn((proto_hw, proto_hw, npz), generator=g))
    # make the quality channel of the confident instances peaky enough to pass threshold_abs = 0.4 after smoothing
    return {"anchors": anchors.view(-1).tolist(), "protos": protos.unsqueeze(0), "cls_pred": torch.softmax(logits, -1).unsqueeze(0),
            "box_pred": box.unsqueeze(0), "ins_coef_pred": coef.unsqueeze(0), "grasp_coef_pred": gcoef.unsqueeze(0)}


# ====================================================================== one seeded GLOBAL batch (multi-GPU J gate)
def make_global_samples(lo: int, hi: int, word_len: int, size: int = 416, max_gt: int = 64, base_seed: int = 9000):
    """Samples [lo, hi) of ONE seeded global batch.  Sample i depends on ``base_seed + i`` only, so every rank of every
    world size produces the same bytes for it: the J counters summed over 1, 2, 4 or 8 shards must be identical
    (shards from crog_b200.engine.shard_range).  Returns (img, word, gt, gt_count)."""
    imgs, words, gts, cnts = [], [], [], []
    for i in range(lo, hi):
        img, word = make_inputs(1, word_len, size, seed_img=base_seed + 3 * i, seed_txt=base_seed + 3 * i + 1)
        gt, cnt = make_gt_rects(1, max_gt, seed=base_seed + 3 * i + 2, size=size)
        imgs.append(img); words.append(word); gts.append(gt); cnts.append(cnt)
    return torch.cat(imgs), torch.cat(words), np.concatenate(gts), np.concatenate(cnts)


def plant_gt_near_predictions(gt: np.ndarray, cnt: np.ndarray, grasps: np.ndarray, n_peaks: np.ndarray, lo: int,
                              base_seed: int = 9000) -> np.ndarray:
    """Random-init weights put the predicted grasps nowhere near random ground truth (J@1 = 0 is a poor parity witness).
    Overwrites GT row 0 of about half the samples with a rectangle near the sample's own top-1 prediction and of another
    quarter near its 3rd prediction (J@5 hit without a J@1 hit), jittered by a per-sample seeded RNG, so that both outcomes
    of both counters occur.  A deterministic function of (sample index, the sample's decoded grasps)."""
    gt = gt.copy()
    for b in range(gt.shape[0]):
        rng = np.random.default_rng(base_seed + 7919 * (lo + b))
        u = rng.random()
        k = 0 if u < 0.5 else (2 if u < 0.75 else -1)
        if k < 0 or n_peaks[b] <= k:
            continue
        x, y, w, _, a = grasps[b, k]
        gt[b, 0] = [x + rng.uniform(-6, 6), y + rng.uniform(-6, 6), min(max(w + rng.uniform(-12, 12), 4.0), 130.0),
                    rng.uniform(10, 40), a + rng.uniform(-22, 22), 1.0]
    return gt


def make_frames_u8(batch: int, h: int = 480, w: int = 640, seed: int = 11) -> torch.Tensor:
    """OCID-VLG sized camera frames: uint8 RGB [B, h, w, 3] (smooth blobs + noise; content is irrelevant to throughput)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand((batch, h // 16, w // 16, 3), generator=g).permute(0, 3, 1, 2)
    up = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    return (up * 200 + torch.rand((batch, h, w, 3), generator=g) * 55).clamp(0, 255).to(torch.uint8).contiguous()
