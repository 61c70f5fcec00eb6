"""Anchor table of the reference SSG (model/ssg.py:229-235, utils/box_utils.py:88-103): input independent."""
from __future__ import annotations

import math
from itertools import product
from typing import List


def make_anchors(cfg, conv_h: int, conv_w: int, scale: float) -> List[float]:
    """utils/box_utils.py:88-103: centre-form priors [x, y, w, h] per cell and aspect ratio, row-major cells."""
    out: List[float] = []
    for j, i in product(range(conv_h), range(conv_w)):
        x = (i + 0.5) / conv_w
        y = (j + 0.5) / conv_h
        for ar in cfg.aspect_ratios:
            ar = math.sqrt(ar)
            out += [x, y, scale * ar / cfg.img_size, scale / ar / cfg.img_size]
    return out


def fpn_shapes(cfg) -> List[int]:
    return [math.ceil(cfg.img_size / s) for s in cfg.anchor_strides]


def make_all_anchors(cfg) -> List[float]:
    """model/ssg.py:229-235."""
    scales = [int(cfg.img_size / 544 * aa) for aa in (24, 48, 96, 192, 384)]
    out: List[float] = []
    for i, size in enumerate(fpn_shapes(cfg)):
        out += make_anchors(cfg, size, size, scales[i])
    return out
