"""Execution plan of the SSG forward (reference model/ssg.py:248-279, eval branch) on the crog_b200 kernels.

Same machinery as the CROG plan (crog_b200/model/plan.py): weights packed once (BN folded, K-major tap-ordered),
every activation an NHWC row matrix (zero-haloed where a 3x3 convolution reads it), launches recorded as a flat list.
What is specific to the torchvision-style trunk:
  * 7x7/2 stem on RGB-D: a patch gather (crog_stem7_patches, K = 49*cin padded to a multiple of 64) + one GEMM;
  * stride-2 3x3 convolutions: crog_patches3 gathers the nine strided taps into a compact [rows, 9C] matrix, so the
    contraction issues exactly the reference FLOPs; stride-2 1x1 convolutions read an x[::2, ::2] copy;
  * the five pyramid levels share the prediction weights; every level's head GEMMs write straight into
    [B, sum_l H_l W_l, N] tensors (out_sample_rows), which IS torch.cat(..., dim=1) of model/ssg.py:266-269.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .. import _lib as L
from .plan import Act, ForwardPlan, _aligned, _bn_fold, _conv_w
from .ssg_anchors import fpn_shapes


class SSGPlan(ForwardPlan):
    def __init__(self, sd: Dict[str, torch.Tensor], cfg, batch: int, precision: str = "bf16",
                 device: Optional[torch.device] = None, gemm_impl: int = L.IMPL_AUTO, keep: bool = False):
        assert precision in ("bf16", "fp32")
        self.lib = L.lib()
        self.cfg, self.B, self.precision = cfg, batch, precision
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.acode = L.dtype_code(self.adt)
        self.impl = gemm_impl
        self.ops, self.op_names, self.keep, self._keep_all, self._hold = [], [], {}, keep, []
        self.op_launches = []
        self.n_launches, self.gemm_flops, self.gemm_alg_flops, self.gemm_alg_bytes = 0, 0, {}, {}
        self.gemm_ops, self.tile_choice, self.gemm_index = [], {}, {}
        self.text_sms, self.front_end, self.stem_pairs = 0, 0, False
        self.snake, self._wdir = os.environ.get("CROG_SNAKE", "1") != "0", {}
        self.op_side, self.op_after, self.side_helpers, self.fuse_downsample = set(), {}, False, False
        self._side = self._ev = None
        self.text_range = (0, 0)
        self.sd = {k: _aligned(v.detach().to(self.dev)) for k, v in sd.items()}
        S = cfg.img_size
        self.S = S
        self.rgb = torch.zeros((batch, 3, S, S), device=self.dev, dtype=torch.float32)
        self.depth = torch.zeros((batch, 1, S, S), device=self.dev, dtype=torch.float32)
        self._build()

    def _random_inputs(self):
        gen = torch.Generator(device="cpu").manual_seed(1234)
        self.rgb.copy_(torch.rand(self.rgb.shape, generator=gen))
        self.depth.copy_(torch.rand(self.depth.shape, generator=gen))

    def _zero_inputs(self):
        self.rgb.zero_()
        self.depth.zero_()

    def run(self, stream: Optional[int] = None, fork_text: bool = False):
        s = stream if stream is not None else L.stream_ptr()
        for fn in self.ops:
            fn(s)

    # ------------------------------------------------------------------ helpers
    def _convb(self, name: str, p: str, a: Act, out: Act, taps: int, act: int = L.ACT_NONE, **kw):
        """conv with bias, no BN (FPN / ProtoNet / heads, model/ssg.py:122-183)."""
        w = self.sd[p + ".weight"]
        self.gemm(name, a, self.wt(_conv_w(w)), w.shape[0], out, taps=taps, bias=self.f32(self.sd[p + ".bias"]), act=act, **kw)

    def _patches(self, name: str, src: Act, stride: int) -> Act:
        OH, OW = (src.H - 1) // stride + 1, (src.W - 1) // stride + 1
        assert src.padded and src.col0 == 0 and src.ld == src.C
        out = self.new(OH, OW, 9 * src.C)
        lib = self.lib
        a = (src.ptr, src.ld, out.ptr, src.B, src.H, src.W, src.C, stride, L.dtype_code(src.t.dtype))
        self._add(name, lambda s: L.check(lib.crog_patches3(*a, s)))
        return out

    def _maxpool(self, name: str, src: Act, dst: Act):
        lib = self.lib
        a = (src.ptr, src.ld, int(src.padded), dst.ptr, dst.ld, int(dst.padded), src.B, src.H, src.W, src.C, L.dtype_code(src.t.dtype))
        self._add(name, lambda s: L.check(lib.crog_maxpool3s2(*a, s)))

    # ------------------------------------------------------------------ the network
    def _build(self):
        sd, cfg, B, S, lib = self.sd, self.cfg, self.B, self.S, self.lib
        RELU = L.ACT_RELU
        # ---- stem (model/ssg.py:65-67,98-101,217-222)
        w1 = sd["backbone.conv1.weight"]
        cin = w1.shape[1]
        Kp = ((49 * cin + 63) // 64) * 64
        H1 = (S - 1) // 2 + 1
        patches = self.new(H1, H1, Kp)
        a = (self.rgb.data_ptr(), self.depth.data_ptr() if cin == 4 else None, B, S, S, cin, Kp, patches.ptr, self.acode)
        self._add("stem.patches", lambda s: L.check(lib.crog_stem7_patches(*a, s)))
        wk = F.pad(w1.permute(0, 2, 3, 1).reshape(64, -1), (0, Kp - 49 * cin))
        sc, bi = _bn_fold(sd, "backbone.bn1")
        s0 = self.new(H1, H1, 64)
        self.gemm("stem.conv1", patches, self.wt(wk), 64, s0, scale=self.f32(sc), bias=self.f32(bi), act=RELU, alg_cin=49 * cin)
        H2 = (H1 - 1) // 2 + 1
        x = self.new(H2, H2, 64)
        self._maxpool("stem.maxpool", s0, x)
        # ---- residual stages (model/ssg.py:15-50,74-95)
        feats: List[Act] = []
        inpl = 64
        for li, nb in enumerate(cfg.resnet_layers):
            planes = 64 * 2 ** li
            for bi_ in range(nb):
                x = self._bottleneck(f"backbone.layers.{li}.{bi_}", x, inpl, planes, 2 if (li > 0 and bi_ == 0) else 1)
                inpl = planes * 4
            feats.append(x)
            self.keep[f"c{li + 2}"] = x
        c3, c4, c5 = feats[1], feats[2], feats[3]
        # ---- FPN (model/ssg.py:189-205)
        p5_1 = self.new(c5.H, c5.W, 256, padded=True)
        self._convb("fpn.lat2", "fpn.lat_layers.2", c5, p5_1, 1)
        up5 = self.new(c4.H, c4.W, 256, padded=True)
        self.resample("fpn.up5", p5_1, up5, L.RS_BILINEAR2)
        p4_1 = self.new(c4.H, c4.W, 256, padded=True)
        self._convb("fpn.lat1", "fpn.lat_layers.1", c4, p4_1, 1, residual=up5)
        up4 = self.new(c3.H, c3.W, 256, padded=True)
        self.resample("fpn.up4", p4_1, up4, L.RS_BILINEAR2)
        p3_1 = self.new(c3.H, c3.W, 256, padded=True)
        self._convb("fpn.lat0", "fpn.lat_layers.0", c3, p3_1, 1, residual=up4)
        pyr: List[Act] = []
        for i, src in ((0, p3_1), (1, p4_1), (2, p5_1)):
            p = self.new(src.H, src.W, 256, padded=True)
            self._convb(f"fpn.pred{i}", f"fpn.pred_layers.{i}.0", src, p, 9, act=RELU)
            pyr.append(p)
        for i in range(2):
            src = pyr[-1]
            pt = self._patches(f"fpn.down{i}.patches", src, 2)
            p = self.new(pt.H, pt.W, 256, padded=True)
            self._convb(f"fpn.down{i}", f"fpn.downsample_layers.{i}.0", pt, p, 1, act=RELU, alg_cin=9 * 256)
            pyr.append(p)
        self.keep.update(p3=pyr[0], p5=pyr[2], p7=pyr[4])
        assert [p.H for p in pyr] == fpn_shapes(cfg), ([p.H for p in pyr], fpn_shapes(cfg))
        # ---- ProtoNet on P3 (model/ssg.py:150-169)
        t = pyr[0]
        for i in (0, 2, 4):
            o = self.new(t.H, t.W, 256, padded=(i != 4))
            self._convb(f"proto1.{i}", f"proto_net.proto1.{i}", t, o, 9, act=RELU)
            t = o
        u = self.new(2 * t.H, 2 * t.W, 256, padded=True)
        self.resample("proto.up", t, u, L.RS_BILINEAR2_AC)
        t = self.new(u.H, u.W, 256)
        self._convb("proto2.0", "proto_net.proto2.0", u, t, 9, act=RELU)
        npz = cfg.num_protos
        self.protos = torch.zeros((B, t.H, t.W, npz), device=self.dev, dtype=torch.float32)
        self._hold.append(self.protos)
        self._convb("proto2.2", "proto_net.proto2.2", t, Act(self.protos.view(-1, npz), B, t.H, t.W, False, npz), 1, act=RELU)
        # ---- prediction heads, shared over the five levels (model/ssg.py:117-147,258-269)
        na, nc = len(cfg.aspect_ratios), cfg.num_classes
        tot = sum(p.H * p.W for p in pyr)
        ncb = ((na * (nc + 4) + 15) // 16) * 16
        self.cb = torch.zeros((B * tot, ncb), device=self.dev, dtype=torch.float32)
        self.coef = torch.zeros((B, tot * na, npz), device=self.dev, dtype=torch.float32)
        self.gcoef = torch.zeros((B, tot * na, 4, npz), device=self.dev, dtype=torch.float32)
        self.cls = torch.zeros((B, tot * na, nc), device=self.dev, dtype=torch.float32)
        self.box = torch.zeros((B, tot * na, 4), device=self.dev, dtype=torch.float32)
        self._hold.extend([self.cb, self.coef, self.gcoef, self.cls, self.box])
        pl = "prediction_layers"
        w_cb = torch.cat([_conv_w(sd[pl + ".conf_layer.weight"]), _conv_w(sd[pl + ".bbox_layer.weight"])])
        w_cb = self.wt(F.pad(w_cb, (0, 0, 0, ncb - w_cb.shape[0])))
        b_cb = self.f32(F.pad(torch.cat([sd[pl + ".conf_layer.bias"], sd[pl + ".bbox_layer.bias"]]).float(), (0, ncb - na * (nc + 4))))
        w_up, b_up = self.wt(_conv_w(sd[pl + ".upfeature.0.weight"])), self.f32(sd[pl + ".upfeature.0.bias"])
        w_co, b_co = self.wt(_conv_w(sd[pl + ".coef_layer.0.weight"])), self.f32(sd[pl + ".coef_layer.0.bias"])
        w_gc, b_gc = self.wt(_conv_w(sd[pl + ".grasp_coef_layer.0.weight"])), self.f32(sd[pl + ".grasp_coef_layer.0.bias"])
        off = 0
        for li, p in enumerate(pyr):
            up = self.new(p.H, p.W, 256, padded=True)
            self.gemm(f"head{li}.upfeature", p, w_up, 256, up, taps=9, bias=b_up, act=L.ACT_RELU)
            kw = dict(taps=9, out_sample_rows=tot, out_row0=off)
            self.gemm(f"head{li}.conf_box", up, w_cb, ncb, Act(self.cb, B, p.H, p.W, False, ncb), bias=b_cb, alg_n=na * (nc + 4), **kw)
            self.gemm(f"head{li}.coef", up, w_co, na * npz, Act(self.coef.view(B * tot, na * npz), B, p.H, p.W, False, na * npz),
                      bias=b_co, act=L.ACT_TANH, **kw)
            self.gemm(f"head{li}.gcoef", up, w_gc, na * npz * 4, Act(self.gcoef.view(B * tot, na * npz * 4), B, p.H, p.W, False, na * npz * 4),
                      bias=b_gc, act=L.ACT_TANH, **kw)
            off += p.H * p.W
        a2 = (self.cb.data_ptr(), ncb, B * tot, na, nc, self.cls.data_ptr(), self.box.data_ptr())
        self._add("heads.softmax", lambda s: L.check(lib.crog_ssg_heads(*a2, s)))

    def _bottleneck(self, p: str, x: Act, inpl: int, planes: int, stride: int) -> Act:
        sd, RELU = self.sd, L.ACT_RELU
        H, W = x.H, x.W
        sc, bi = _bn_fold(sd, p + ".bn1")
        t1 = self.new(H, W, planes, padded=True)
        self.gemm(p + ".conv1", x, self.wt(_conv_w(sd[p + ".conv1.weight"])), planes, t1, scale=self.f32(sc), bias=self.f32(bi), act=RELU)
        sc, bi = _bn_fold(sd, p + ".bn2")
        w2 = self.wt(_conv_w(sd[p + ".conv2.weight"]))
        if stride == 1:
            t2 = self.new(H, W, planes)
            self.gemm(p + ".conv2", t1, w2, planes, t2, taps=9, scale=self.f32(sc), bias=self.f32(bi), act=RELU)
        else:
            pt = self._patches(p + ".conv2.patches", t1, stride)
            t2 = self.new(pt.H, pt.W, planes)
            self.gemm(p + ".conv2", pt, w2, planes, t2, scale=self.f32(sc), bias=self.f32(bi), act=RELU)
        idt = x
        if (p + ".downsample.0.weight") in sd:
            xi = x
            if stride > 1:
                xi = self.new(t2.H, t2.W, inpl)
                self.resample(p + ".downsample.sub", x, xi, L.RS_SUBSAMPLE2)
            sc, bi = _bn_fold(sd, p + ".downsample.1")
            idt = self.new(xi.H, xi.W, planes * 4)
            self.gemm(p + ".downsample", xi, self.wt(_conv_w(sd[p + ".downsample.0.weight"])), planes * 4, idt, scale=self.f32(sc), bias=self.f32(bi))
        sc, bi = _bn_fold(sd, p + ".bn3")
        out = self.new(t2.H, t2.W, planes * 4)
        self.gemm(p + ".conv3", t2, self.wt(_conv_w(sd[p + ".conv3.weight"])), planes * 4, out, scale=self.f32(sc), bias=self.f32(bi),
                  residual=idt, residual_relu=True)
        return out
