"""Tensor inventory of the reference CROG module, as a flat table.

The drop-in contract is that ``state_dict()`` of our module has
exactly the reference's 662 names and shapes, so a reference checkpoint loads with
``strict=True``.  Instead of mirroring the reference's class tree we generate the
table of (dotted name, shape, role) once and build both the module skeleton
(``crog_b200.model.crog``) and synthetic weights (``crog_b200.synth``) from it.

Reference for the names/shapes: ``model/clip.py:10-57,60-78,157-199,239-283,334-376``
(CLIP RN50 towers), ``model/layers.py:47-62,176-194,280-312,342-370`` (neck, decoder,
projector) and ``model/crog.py:20-45`` (top-level attribute names).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple

# CLIP RN50 hyper-parameters that build_model (model/clip.py:503-546) infers from
# the RN50 checkpoint.
CLIP_EMBED_DIM = 1024
CLIP_VISION_LAYERS = (3, 4, 6, 3)
CLIP_VISION_WIDTH = 64
CLIP_CONTEXT_LEN = 77
CLIP_VOCAB = 49408
CLIP_TXT_WIDTH = 512
CLIP_TXT_HEADS = 8
CLIP_TXT_LAYERS = 12
CLIP_POOL_GRID = 7  # image_resolution 224 // 32


@dataclass(frozen=True)
class TensorSpec:
    name: str
    shape: Tuple[int, ...]
    role: str  # conv | linear_w | bias | bn_w | bn_b | bn_mean | bn_var | bn_count |
    #            ln_w | ln_b | embed | pos | proj | scalar
    fan_in: int = 0
    is_buffer: bool = False
    dtype: str = "float32"


def _bn(prefix: str, ch: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (ch,), "bn_w")
    yield TensorSpec(prefix + ".bias", (ch,), "bn_b")
    yield TensorSpec(prefix + ".running_mean", (ch,), "bn_mean", is_buffer=True)
    yield TensorSpec(prefix + ".running_var", (ch,), "bn_var", is_buffer=True)
    yield TensorSpec(prefix + ".num_batches_tracked", (), "bn_count", is_buffer=True, dtype="int64")


def _conv(name: str, cout: int, cin: int, k: int) -> TensorSpec:
    return TensorSpec(name, (cout, cin, k, k), "conv", fan_in=cin * k * k)


def _linear(prefix: str, cout: int, cin: int, bias: bool = True) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (cout, cin), "linear_w", fan_in=cin)
    if bias:
        yield TensorSpec(prefix + ".bias", (cout,), "bias", fan_in=cin)


def _ln(prefix: str, ch: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (ch,), "ln_w")
    yield TensorSpec(prefix + ".bias", (ch,), "ln_b")


def _mha(prefix: str, d: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".in_proj_weight", (3 * d, d), "linear_w", fan_in=d)
    yield TensorSpec(prefix + ".in_proj_bias", (3 * d,), "bias", fan_in=d)
    yield from _linear(prefix + ".out_proj", d, d)


def _cbr(prefix: str, cout: int, cin: int, k: int) -> Iterator[TensorSpec]:
    """conv(no bias) + BN as nn.Sequential indices .0 / .1 (layers.py:8-11)."""
    yield _conv(prefix + ".0.weight", cout, cin, k)
    yield from _bn(prefix + ".1", cout)


def crog_tensor_specs(cfg) -> List[TensorSpec]:
    """All tensors of ``CROG(cfg).state_dict()`` for the RN50 / MultiTaskProjector
    configuration (use_contrastive and use_grasp_masks both on or off per cfg)."""
    out: List[TensorSpec] = []
    add = out.append
    ext = out.extend
    w = CLIP_VISION_WIDTH

    # ---- backbone: CLIP top-level (clip.py:366-376)
    add(TensorSpec("backbone.positional_embedding", (CLIP_CONTEXT_LEN, CLIP_TXT_WIDTH), "pos"))
    add(TensorSpec("backbone.text_projection", (CLIP_TXT_WIDTH, CLIP_EMBED_DIM), "proj", fan_in=CLIP_TXT_WIDTH))
    add(TensorSpec("backbone.logit_scale", (), "scalar"))
    add(TensorSpec("backbone.token_embedding.weight", (CLIP_VOCAB, CLIP_TXT_WIDTH), "embed"))
    ext(_ln("backbone.ln_final", CLIP_TXT_WIDTH))

    # ---- backbone.visual: stem (clip.py:165-184)
    v = "backbone.visual"
    add(_conv(v + ".conv1.weight", w // 2, 3, 3)); ext(_bn(v + ".bn1", w // 2))
    add(_conv(v + ".conv2.weight", w // 2, w // 2, 3)); ext(_bn(v + ".bn2", w // 2))
    add(_conv(v + ".conv3.weight", w, w // 2, 3)); ext(_bn(v + ".bn3", w))
    # residual stages (clip.py:187-203, 10-42)
    inplanes = w
    for li, nblocks in enumerate(CLIP_VISION_LAYERS, start=1):
        planes = w * (2 ** (li - 1))
        for bi in range(nblocks):
            p = f"{v}.layer{li}.{bi}"
            add(_conv(p + ".conv1.weight", planes, inplanes, 1)); ext(_bn(p + ".bn1", planes))
            add(_conv(p + ".conv2.weight", planes, planes, 3)); ext(_bn(p + ".bn2", planes))
            add(_conv(p + ".conv3.weight", planes * 4, planes, 1)); ext(_bn(p + ".bn3", planes * 4))
            if bi == 0:
                add(_conv(p + ".downsample.0.weight", planes * 4, inplanes, 1))
                ext(_bn(p + ".downsample.1", planes * 4))
            inplanes = planes * 4
    # attention pool (clip.py:60-78)
    e = w * 32
    a = v + ".attnpool"
    add(TensorSpec(a + ".positional_embedding", (CLIP_POOL_GRID ** 2 + 1, e), "pos"))
    for nm in ("k_proj", "q_proj", "v_proj"):
        ext(_linear(f"{a}.{nm}", e, e))
    ext(_linear(a + ".c_proj", CLIP_EMBED_DIM, e))
    ext(_cbr(a + ".connect", CLIP_EMBED_DIM, e, 1))

    # ---- backbone.transformer (clip.py:239-283)
    for i in range(CLIP_TXT_LAYERS):
        p = f"backbone.transformer.resblocks.{i}"
        ext(_mha(p + ".attn", CLIP_TXT_WIDTH))
        ext(_ln(p + ".ln_1", CLIP_TXT_WIDTH))
        ext(_linear(p + ".mlp.c_fc", 4 * CLIP_TXT_WIDTH, CLIP_TXT_WIDTH))
        ext(_linear(p + ".mlp.c_proj", CLIP_TXT_WIDTH, 4 * CLIP_TXT_WIDTH))
        ext(_ln(p + ".ln_2", CLIP_TXT_WIDTH))

    # ---- neck (layers.py:342-370)
    fi, fo = list(cfg.fpn_in), list(cfg.fpn_out)
    ext(_linear("neck.txt_proj.0", fo[2], fi[2], bias=False)); ext(_bn("neck.txt_proj.1", fo[2]))
    ext(_cbr("neck.f1_v_proj", fo[2], fi[2], 1))
    ext(_bn("neck.norm_layer.0", fo[2]))
    ext(_cbr("neck.f2_v_proj", fo[1], fi[1], 3))
    ext(_cbr("neck.f2_cat", fo[1], fo[2] + fo[1], 1))
    ext(_cbr("neck.f3_v_proj", fo[0], fi[0], 3))
    ext(_cbr("neck.f3_cat", fo[1], fo[0] + fo[1], 1))
    ext(_cbr("neck.f4_proj5", fo[1], fo[2], 3))
    ext(_cbr("neck.f4_proj4", fo[1], fo[1], 3))
    ext(_cbr("neck.f4_proj3", fo[1], fo[1], 3))
    ext(_cbr("neck.aggr", fo[1], 3 * fo[1], 1))
    ext(_cbr("neck.coordconv.0.conv1", fo[1], fo[1] + 2, 3))
    ext(_cbr("neck.coordconv.1", fo[1], fo[1], 3))

    # ---- decoder (layers.py:176-194, 280-312)
    if cfg.use_contrastive:
        d, f = cfg.vis_dim, cfg.dim_ffn
        for i in range(cfg.num_layers):
            p = f"decoder.layers.{i}"
            ext(_ln(p + ".self_attn_norm", d)); ext(_ln(p + ".cross_attn_norm", d))
            ext(_mha(p + ".self_attn", d)); ext(_mha(p + ".multihead_attn", d))
            ext(_linear(p + ".ffn.0", f, d)); ext(_ln(p + ".ffn.3", f)); ext(_linear(p + ".ffn.4", d, f))
            ext(_ln(p + ".norm1", d)); ext(_ln(p + ".norm2", d)); ext(_ln(p + ".norm3", d))
        ext(_ln("decoder.norm", cfg.vis_dim))

    # ---- projector (layers.py:47-62 / 135-150)
    c = cfg.vis_dim // 2
    heads = 5 if cfg.use_grasp_masks else 1
    ext(_cbr("proj.vis.1", 2 * c, 2 * c, 3))
    ext(_cbr("proj.vis.3", c, 2 * c, 3))
    add(_conv("proj.vis.4.weight", c * heads, c, 1))
    add(TensorSpec("proj.vis.4.bias", (c * heads,), "bias", fan_in=c))
    ext(_linear("proj.txt", c * 9 + 1, cfg.word_dim))
    return out


# ====================================================================== SSG (config 4)
def ssg_tensor_specs(cfg) -> List[TensorSpec]:
    """All tensors of the reference ``SSG(cfg).state_dict()`` (model/ssg.py:53-245): torchvision-style
    ResNet (7x7 stem on 3 or 4 channels, stride in the 3x3 conv), FPN with biased convs and no BN,
    ProtoNet, the shared PredictionModule and ``semantic_seg_conv`` (created because the module is
    built in training mode, ssg.py:237-238; unused at inference but part of every checkpoint)."""
    out: List[TensorSpec] = []
    add, ext = out.append, out.extend

    def convb(name: str, cout: int, cin: int, k: int):
        add(_conv(name + ".weight", cout, cin, k))
        add(TensorSpec(name + ".bias", (cout,), "bias", fan_in=cin * k * k))

    cin0 = 4 if cfg.with_depth else 3
    add(_conv("backbone.conv1.weight", 64, cin0, 7)); ext(_bn("backbone.bn1", 64))
    inplanes = 64
    for li, nblocks in enumerate(cfg.resnet_layers):
        planes = 64 * 2 ** li
        for bi in range(nblocks):
            p = f"backbone.layers.{li}.{bi}"
            add(_conv(p + ".conv1.weight", planes, inplanes, 1)); ext(_bn(p + ".bn1", planes))
            add(_conv(p + ".conv2.weight", planes, planes, 3)); ext(_bn(p + ".bn2", planes))
            add(_conv(p + ".conv3.weight", planes * 4, planes, 1)); ext(_bn(p + ".bn3", planes * 4))
            if bi == 0:
                add(_conv(p + ".downsample.0.weight", planes * 4, inplanes, 1))
                ext(_bn(p + ".downsample.1", planes * 4))
            inplanes = planes * 4
    for i, c in enumerate(cfg.fpn_in_channels):
        convb(f"fpn.lat_layers.{i}", 256, c, 1)
    for i in range(len(cfg.fpn_in_channels)):
        convb(f"fpn.pred_layers.{i}.0", 256, 256, 3)
    for i in range(2):
        convb(f"fpn.downsample_layers.{i}.0", 256, 256, 3)
    for i in (0, 2, 4):
        convb(f"proto_net.proto1.{i}", 256, 256, 3)
    convb("proto_net.proto2.0", 256, 256, 3)
    convb("proto_net.proto2.2", cfg.num_protos, 256, 1)
    na = len(cfg.aspect_ratios)
    convb("prediction_layers.upfeature.0", 256, 256, 3)
    convb("prediction_layers.bbox_layer", na * 4, 256, 3)
    convb("prediction_layers.conf_layer", na * cfg.num_classes, 256, 3)
    convb("prediction_layers.coef_layer.0", na * cfg.num_protos, 256, 3)
    if cfg.with_grasp_masks:
        convb("prediction_layers.grasp_coef_layer.0", na * cfg.num_protos * 4, 256, 3)
    convb("semantic_seg_conv", cfg.num_classes, 256, 1)
    return out
