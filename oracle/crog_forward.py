"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the reference CROG inference forward as one flat function
over a state-dict.  It is written against plain ``torch.nn.functional`` ops and is
pinned two ways:
  * ``oracle/make_golden.py`` (run in the build container) imports the real
    reference from /root/reference, loads the same synthetic state-dict and
    asserts this restatement reproduces it, then writes tests/golden/*.npz;
  * ``tests/test_oracle_model.py`` re-checks the restatement against those
    committed golden vectors wherever the repo is checked out.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may
import this file.

Reference lines restated:
  model/crog.py:47-113           forward orchestration, pad mask
  model/clip.py:44-57            Bottleneck
  model/clip.py:80-144           AttentionPool2d (bicubic pos-embed, connect residual)
  model/clip.py:207-223          ModifiedResNet.forward
  model/clip.py:239-265,439-456  text transformer, encode_text
  model/layers.py:371-398        FPN.forward
  model/layers.py:195-277,313-339 decoder
  model/layers.py:64-132         MultiTaskProjector.forward
  engine/crog_engine.py:183-211  sigmoid + bicubic x4 glue
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

EPS_BN = 1e-5
EPS_LN = 1e-5


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, EPS_BN)


def _cbr(sd, p, x, pad):
    """conv(no bias) + BN + ReLU, nn.Sequential indices 0/1 (layers.py:8-11)."""
    return F.relu(_bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"], padding=pad)))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], EPS_LN)


def _mha(q, k, v, w_in, b_in, w_out, b_out, heads, attn_mask=None, key_padding_mask=None):
    """Packed-weight multi-head attention on batch-first tensors [B,T,D]
    (== nn.MultiheadAttention eval semantics; q scaled by head_dim**-0.5)."""
    B, Tq, D = q.shape
    Tk = k.shape[1]
    hd = D // heads
    wq, wk, wv = w_in.chunk(3, 0)
    bq, bk, bv = b_in.chunk(3, 0)
    Q = (F.linear(q, wq, bq) * hd ** -0.5).view(B, Tq, heads, hd).transpose(1, 2)
    K = F.linear(k, wk, bk).view(B, Tk, heads, hd).transpose(1, 2)
    V = F.linear(v, wv, bv).view(B, Tk, heads, hd).transpose(1, 2)
    s = Q @ K.transpose(-1, -2)
    if attn_mask is not None:
        s = s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, -1) @ V
    o = o.transpose(1, 2).reshape(B, Tq, D)
    return F.linear(o, w_out, b_out)


def _bottleneck(sd, p, x, stride):
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], padding=1)))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    idt = x
    if (p + ".downsample.0.weight") in sd:
        idt = F.avg_pool2d(x, stride) if stride > 1 else x
        idt = _bn(sd, p + ".downsample.1", F.conv2d(idt, sd[p + ".downsample.0.weight"]))
    return F.relu(out + idt)


def encode_image(sd, img, inter=None):
    v = "backbone.visual"
    x = F.relu(_bn(sd, v + ".bn1", F.conv2d(img, sd[v + ".conv1.weight"], stride=2, padding=1)))
    x = F.relu(_bn(sd, v + ".bn2", F.conv2d(x, sd[v + ".conv2.weight"], padding=1)))
    x = F.relu(_bn(sd, v + ".bn3", F.conv2d(x, sd[v + ".conv3.weight"], padding=1)))
    x = F.avg_pool2d(x, 2)
    if inter is not None:
        inter["stem"] = x
    feats = []
    for li, nb in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nb):
            x = _bottleneck(sd, f"{v}.layer{li}.{bi}", x, 2 if (li > 1 and bi == 0) else 1)
        feats.append(x)
        if inter is not None:
            inter[f"layer{li}"] = x
    x4 = feats[3]
    a = v + ".attnpool"
    B, C, H, W = x4.shape
    res = _bn(sd, a + ".connect.1", F.conv2d(x4, sd[a + ".connect.0.weight"]))
    pe = sd[a + ".positional_embedding"]
    g = int(round(math.sqrt(pe.shape[0] - 1)))
    pe = pe[1:].reshape(1, g, g, C).permute(0, 3, 1, 2)
    pe = F.interpolate(pe, size=(H, W), mode="bicubic", align_corners=False)
    tok = (x4 + pe).flatten(2).transpose(1, 2)  # B, HW, C
    w_in = torch.cat([sd[a + ".q_proj.weight"], sd[a + ".k_proj.weight"], sd[a + ".v_proj.weight"]])
    b_in = torch.cat([sd[a + ".q_proj.bias"], sd[a + ".k_proj.bias"], sd[a + ".v_proj.bias"]])
    # output dim (1024) != embed dim (2048): do the projections by hand
    heads = C // 64
    hd = 64
    wq, wk, wv = w_in.chunk(3, 0)
    bq, bk, bv = b_in.chunk(3, 0)
    Q = (F.linear(tok, wq, bq) * hd ** -0.5).view(B, H * W, heads, hd).transpose(1, 2)
    K = F.linear(tok, wk, bk).view(B, H * W, heads, hd).transpose(1, 2)
    V = F.linear(tok, wv, bv).view(B, H * W, heads, hd).transpose(1, 2)
    o = torch.softmax(Q @ K.transpose(-1, -2), -1) @ V
    o = o.transpose(1, 2).reshape(B, H * W, C)
    o = F.linear(o, sd[a + ".c_proj.weight"], sd[a + ".c_proj.bias"])
    o = o.transpose(1, 2).reshape(B, -1, H, W)
    c5 = F.relu(o + res)
    return feats[1], feats[2], c5


def encode_text(sd, word, heads=8):
    B, L = word.shape
    x = sd["backbone.token_embedding.weight"][word] + sd["backbone.positional_embedding"][:L]
    causal = torch.full((L, L), float("-inf")).triu_(1)
    i = 0
    while f"backbone.transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"backbone.transformer.resblocks.{i}"
        h = _ln(sd, p + ".ln_1", x)
        x = x + _mha(h, h, h, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"],
                     sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], heads, attn_mask=causal)
        h = _ln(sd, p + ".ln_2", x)
        h = F.linear(h, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + F.linear(h, sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])
        i += 1
    x = _ln(sd, "backbone.ln_final", x)
    state = x[torch.arange(B), word.argmax(-1)] @ sd["backbone.text_projection"]
    return x, state


def neck(sd, c3, c4, c5, state):
    s = F.linear(state, sd["neck.txt_proj.0.weight"])
    s = F.relu(F.batch_norm(s, sd["neck.txt_proj.1.running_mean"], sd["neck.txt_proj.1.running_var"],
                            sd["neck.txt_proj.1.weight"], sd["neck.txt_proj.1.bias"], False, 0.0, EPS_BN))
    f5 = _cbr(sd, "neck.f1_v_proj", c5, 0)
    f5 = F.relu(_bn(sd, "neck.norm_layer.0", f5 * s[:, :, None, None]))
    f4 = _cbr(sd, "neck.f2_v_proj", c4, 1)
    f5u = F.interpolate(f5, scale_factor=2, mode="bilinear")
    f4 = _cbr(sd, "neck.f2_cat", torch.cat([f4, f5u], 1), 0)
    f3 = F.avg_pool2d(_cbr(sd, "neck.f3_v_proj", c3, 1), 2, 2)
    f3 = _cbr(sd, "neck.f3_cat", torch.cat([f3, f4], 1), 0)
    fq5 = F.interpolate(_cbr(sd, "neck.f4_proj5", f5, 1), scale_factor=2, mode="bilinear")
    fq4 = _cbr(sd, "neck.f4_proj4", f4, 1)
    fq3 = _cbr(sd, "neck.f4_proj3", f3, 1)
    fq = _cbr(sd, "neck.aggr", torch.cat([fq3, fq4, fq5], 1), 0)
    B, _, H, W = fq.shape
    ys = torch.linspace(-1, 1, H).view(1, 1, H, 1).expand(B, 1, H, W)
    xs = torch.linspace(-1, 1, W).view(1, 1, 1, W).expand(B, 1, H, W)
    fq = _cbr(sd, "neck.coordconv.0.conv1", torch.cat([fq, xs, ys], 1), 1)
    fq = _cbr(sd, "neck.coordconv.1", fq, 1)
    return fq


def pos2d(d_model, H, W):
    """layers.py:216-241, returned as [H*W, d_model]."""
    pe = torch.zeros(d_model, H, W)
    half = d_model // 2
    div = torch.exp(torch.arange(0.0, half, 2) * -(math.log(10000.0) / half))
    pw = torch.arange(0.0, W).unsqueeze(1) * div  # W, half/2
    ph = torch.arange(0.0, H).unsqueeze(1) * div
    pe[0:half:2] = torch.sin(pw).t().unsqueeze(1).expand(-1, H, -1)
    pe[1:half:2] = torch.cos(pw).t().unsqueeze(1).expand(-1, H, -1)
    pe[half::2] = torch.sin(ph).t().unsqueeze(2).expand(-1, -1, W)
    pe[half + 1::2] = torch.cos(ph).t().unsqueeze(2).expand(-1, -1, W)
    return pe.reshape(d_model, H * W).t().contiguous()


def pos1d(d_model, L):
    """layers.py:195-213, returned as [L, d_model]."""
    pe = torch.zeros(L, d_model)
    position = torch.arange(0, L).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div)
    pe[:, 1::2] = torch.cos(position * div)
    return pe


def decoder(sd, fq, word, pad_mask, heads):
    B, C, H, W = fq.shape
    vis = fq.flatten(2).transpose(1, 2)  # B, HW, C
    vp = pos2d(C, H, W)[None]
    tp = pos1d(word.shape[-1], word.shape[1])[None]
    i = 0
    while f"decoder.layers.{i}.norm1.weight" in sd:
        p = f"decoder.layers.{i}"
        v2 = _ln(sd, p + ".norm1", vis)
        qk = v2 + vp
        v2 = _mha(qk, qk, v2, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"],
                  sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"], heads)
        vis = vis + _ln(sd, p + ".self_attn_norm", v2)
        v2 = _ln(sd, p + ".norm2", vis)
        v2 = _mha(v2 + vp, word + tp, word, sd[p + ".multihead_attn.in_proj_weight"],
                  sd[p + ".multihead_attn.in_proj_bias"], sd[p + ".multihead_attn.out_proj.weight"],
                  sd[p + ".multihead_attn.out_proj.bias"], heads, key_padding_mask=pad_mask)
        vis = vis + _ln(sd, p + ".cross_attn_norm", v2)
        v2 = _ln(sd, p + ".norm3", vis)
        v2 = F.relu(F.linear(v2, sd[p + ".ffn.0.weight"], sd[p + ".ffn.0.bias"]))
        v2 = _ln(sd, p + ".ffn.3", v2)
        vis = vis + F.linear(v2, sd[p + ".ffn.4.weight"], sd[p + ".ffn.4.bias"])
        i += 1
    vis = _ln(sd, "decoder.norm", vis)
    return vis.transpose(1, 2).reshape(B, C, H, W)


def projector(sd, fq, state):
    x = F.interpolate(fq, scale_factor=2, mode="bilinear")
    x = _cbr(sd, "proj.vis.1", x, 1)
    x = F.interpolate(x, scale_factor=2, mode="bilinear")
    x = _cbr(sd, "proj.vis.3", x, 1)
    x = F.conv2d(x, sd["proj.vis.4.weight"], sd["proj.vis.4.bias"])
    B, C5, H, W = x.shape
    nh = C5 // (sd["proj.vis.3.0.weight"].shape[0])
    C = C5 // nh
    t = F.linear(state, sd["proj.txt.weight"], sd["proj.txt.bias"])
    wdyn, bdyn = t[:, :-1].reshape(B, C, 3, 3), t[:, -1]
    outs = []
    for h in range(nh):
        xh = x[:, h * C:(h + 1) * C].reshape(1, B * C, H, W)
        outs.append(F.conv2d(xh, wdyn, bdyn, padding=1, groups=B).transpose(0, 1))
    return outs


@torch.no_grad()
def crog_forward(sd: Dict[str, torch.Tensor], cfg, img: torch.Tensor, word: torch.Tensor,
                 keep: bool = False):
    """Returns (maps, inter): maps = 5 logits B×1×104×104 (mask, qua, sin, cos, wid)."""
    inter = {} if keep else None
    pad = word == 0
    c3, c4, c5 = encode_image(sd, img.float(), inter)
    wfeat, state = encode_text(sd, word)
    fq = neck(sd, c3, c4, c5, state)
    if keep:
        inter.update(c3=c3, c4=c4, c5=c5, word=wfeat, state=state, fq_neck=fq)
    if cfg.use_contrastive:
        fq = decoder(sd, fq, wfeat, pad, cfg.num_head)
        if keep:
            inter["fq_dec"] = fq
    maps = projector(sd, fq, state)
    return maps, inter


@torch.no_grad()
def postprocess(maps, size):
    """engine/crog_engine.py:183-211: sigmoid on mask/qua/wid, bicubic (align_corners=True)
    of all five to ``size``; returns 5 tensors B×H×W."""
    out = []
    for i, m in enumerate(maps):
        if i in (0, 1, 4):
            m = torch.sigmoid(m)
        if tuple(m.shape[-2:]) != tuple(size):
            m = F.interpolate(m, size=size, mode="bicubic", align_corners=True)
        out.append(m.squeeze(1))
    return out
