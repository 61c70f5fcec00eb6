"""Drop-in for the reference's ``model.ssg.SSG`` (model/ssg.py:208-293), inference only (BASELINE config 4).

Same constructor keys (``backbone, resnet_layers, with_depth, fpn_in_channels, num_protos, num_classes,
aspect_ratios, anchor_strides, with_grasp_masks, img_size``), same ``forward(data_dict)`` taking
``{"rgb": B x 3 x S x S, "depth": B x 1 x S x S}`` and returning the reference's eval-mode ``output_dict``
(``anchors`` list, ``protos`` B x S/4 x S/4 x 32, ``cls_pred`` (softmaxed), ``box_pred``, ``ins_coef_pred``,
``grasp_coef_pred``), same ``state_dict()`` names / shapes (a reference checkpoint loads with ``strict=True``).
The forward is an execution plan of hand-written sm_100a kernels (crog_b200/model/ssg_plan.py) replayed as a CUDA
graph; there is no CPU path.  The training branch (losses, model/ssg.py:281-529) is out of scope.
"""
from __future__ import annotations

import os

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib as L
from ..spec import ssg_tensor_specs
from .crog import _capture_stream, _holder
from .ssg_anchors import make_all_anchors


class SSG(nn.Module):
    def __init__(self, cfg, precision: Optional[str] = None, use_cuda_graph: bool = True):
        super().__init__()
        if cfg.backbone != "resnet":
            raise NotImplementedError(cfg.backbone)  # model/ssg.py:225-226
        self.cfg = cfg
        self.precision = precision or getattr(cfg, "precision", "bf16")
        self.use_cuda_graph = use_cuda_graph
        self.gemm_impl = L.IMPL_AUTO
        self.anchors = make_all_anchors(cfg)  # model/ssg.py:229-235
        gen = torch.Generator().manual_seed(0)
        for s in ssg_tensor_specs(cfg):
            mod, leaf = _holder(self, s.name)
            if s.dtype == "int64":
                t = torch.zeros(s.shape, dtype=torch.int64)
            elif s.role in ("bn_w", "bn_var"):
                t = torch.ones(s.shape)
            elif s.role in ("bn_b", "bn_mean", "bias"):
                t = torch.zeros(s.shape)
            else:  # xavier-uniform like the reference (model/ssg.py:241-245)
                co, ci, kh, kw = s.shape
                bound = (6.0 / ((ci + co) * kh * kw)) ** 0.5
                t = (torch.rand(s.shape, generator=gen) * 2 - 1) * bound
            if s.is_buffer:
                mod.register_buffer(leaf, t)
            else:
                mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
        self._plans: Dict[Tuple, object] = {}
        self._graphs: Dict[int, object] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())
        self.eval()

    def invalidate(self):
        self._plans.clear()
        self._graphs.clear()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self.invalidate()
        return super()._apply(fn, *a, **kw)

    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        if precision != self.precision:
            self.precision = precision
            self.invalidate()
        return self

    def plan_for(self, batch: int, keep: bool = False):
        from .ssg_plan import SSGPlan

        dev = self.backbone.conv1.weight.device
        if dev.type != "cuda":
            raise L.CrogError("crog_b200.SSG.forward needs the module on a CUDA device (sm_100a); no CPU fallback exists")
        key = (batch, self.precision, self.gemm_impl, keep, dev.index)
        plan = self._plans.get(key)
        if plan is None:
            with torch.cuda.device(dev):
                plan = SSGPlan(self.state_dict(), self.cfg, batch, self.precision, dev, self.gemm_impl, keep)
                if os.environ.get("CROG_AUTOTUNE", "1") != "0":
                    plan.autotune()  # per-layer tcgen05 tile choice (bf16 plans only; results are unchanged)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError("crog_b200.SSG implements the inference path only; call .eval()")
        rgb = data_dict["rgb"]
        S = self.cfg.img_size
        if rgb.dim() != 4 or rgb.shape[1] != 3 or rgb.shape[2] != S or rgb.shape[3] != S:
            raise RuntimeError(f"expected rgb of shape [B,3,{S},{S}], got {tuple(rgb.shape)}")
        B = rgb.shape[0]
        plan = self.plan_for(B)
        with torch.cuda.device(plan.dev):
            plan.rgb.copy_(rgb, non_blocking=True)
            if self.cfg.with_depth:
                depth = data_dict["depth"]
                if tuple(depth.shape) != (B, 1, S, S):
                    raise RuntimeError(f"expected depth of shape [{B},1,{S},{S}], got {tuple(depth.shape)}")
                plan.depth.copy_(depth, non_blocking=True)
            self._run(plan)
            out = {"anchors": self.anchors, "protos": plan.protos.clone(), "cls_pred": plan.cls.clone(), "box_pred": plan.box.clone(),
                   "ins_coef_pred": plan.coef.clone(), "grasp_coef_pred": plan.gcoef.clone()}
        return out

    def _run(self, plan):
        if not self.use_cuda_graph:
            plan.run()
            return
        g = self._graphs.get(id(plan))
        if g is None:
            plan.run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=_capture_stream(plan.dev)):
                plan.run()
            self._graphs[id(plan)] = g
        g.replay()
