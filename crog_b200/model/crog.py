"""Drop-in for the reference's ``model.crog.CROG`` (model/crog.py:10-133), inference only.

Same constructor (reads the same cfg keys), same ``forward(img, word, mask=None, ...)``
signature and eval-mode return tuple, same ``state_dict()`` names and shapes (so a reference
checkpoint loads with ``strict=True``, with or without the DataParallel ``module.`` prefix) —
but the forward is an execution plan of hand-written sm_100a kernels (crog_b200/model/plan.py),
replayed as a CUDA graph.  There is no CPU path: calling forward without a B200 raises.

Differences by design: training mode is not implemented (the reference's loss branch,
model/crog.py:76-111, is out of scope); ``clip_pretrain`` is not read here — weights arrive
through ``load_state_dict`` (random init otherwise).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from ..spec import crog_tensor_specs
from .. import _lib as L


def _holder(root: nn.Module, dotted: str) -> Tuple[nn.Module, str]:
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, nn.Module())
        m = m._modules[p]
    return m, parts[-1]


class CROG(nn.Module):
    def __init__(self, cfg, precision: Optional[str] = None, use_cuda_graph: bool = True):
        super().__init__()
        self.cfg = cfg
        self.use_contrastive = cfg.use_contrastive
        self.use_pretrained_clip = cfg.use_pretrained_clip
        self.use_grasp_masks = cfg.use_grasp_masks
        self.precision = precision or getattr(cfg, "precision", "bf16")
        self.use_cuda_graph = use_cuda_graph
        self.gemm_impl = L.IMPL_AUTO
        self.autotune = True  # time the tcgen05 tile configurations per layer when a plan is built (CROG_AUTOTUNE=0: off)
        gen = torch.Generator().manual_seed(0)
        for s in crog_tensor_specs(cfg):
            mod, leaf = _holder(self, s.name)
            if s.dtype == "int64":
                t = torch.zeros(s.shape, dtype=torch.int64)
            elif s.role in ("bn_w", "bn_var", "ln_w"):
                t = torch.ones(s.shape)
            elif s.role in ("bn_b", "bn_mean", "ln_b", "bias"):
                t = torch.zeros(s.shape)
            elif s.role == "scalar":
                t = torch.tensor(2.6592600)
            else:
                t = torch.randn(s.shape, generator=gen) * (max(s.fan_in, 1) ** -0.5 if s.fan_in else 0.02)
            if s.is_buffer:
                mod.register_buffer(leaf, t)
            else:
                mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
        self._plans: Dict[Tuple, object] = {}
        self._graphs: Dict[Tuple, object] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())
        self.eval()

    # ------------------------------------------------------------------ weight / plan management
    def invalidate(self):
        """Drop packed weights and captured graphs (call after changing parameters in place)."""
        self._plans.clear()
        self._graphs.clear()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # reference checkpoints are saved from DataParallel/DDP (test_crog.py:70,79): accept "module." keys
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def load_clip(self, clip, strict_shapes: bool = True):
        """Initialise the backbone from a CLIP RN50 archive path (TorchScript ``RN50.pt``) or CLIP state-dict, with the
        fp16 round trip of the reference's ``build_model(..., load_weights=True).float()`` (model/crog.py:20-23;
        crog_b200/model/clip_ingest.py).  Returns (missing backbone names, unexpected archive names)."""
        from .clip_ingest import clip_state_to_backbone, load_clip_archive

        sd = load_clip_archive(clip) if isinstance(clip, (str, bytes)) or hasattr(clip, "__fspath__") else clip
        new, missing, unexpected = clip_state_to_backbone(sd, self.cfg)
        super().load_state_dict(new, strict=False)
        self.invalidate()
        return missing, unexpected

    def _apply(self, fn, *a, **kw):
        self.invalidate()
        return super()._apply(fn, *a, **kw)

    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        if precision != self.precision:
            self.precision = precision
            self.invalidate()
        return self

    def plan_for(self, batch: int, size: int, keep: bool = False):
        from .plan import ForwardPlan

        dev = self.backbone.text_projection.device
        if dev.type != "cuda":
            raise L.CrogError("crog_b200.CROG.forward needs the module on a CUDA device (sm_100a); no CPU fallback exists")
        key = (batch, size, self.precision, self.gemm_impl, keep, dev.index)
        plan = self._plans.get(key)
        if plan is None:
            with torch.cuda.device(dev):
                plan = ForwardPlan(self.state_dict(), self.cfg, batch, self.precision, dev, size, self.gemm_impl, keep)
                if self.autotune and os.environ.get("CROG_AUTOTUNE", "1") != "0":
                    plan.autotune()
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, img, word, mask=None, grasp_qua_mask=None, grasp_sin_mask=None, grasp_cos_mask=None,
                grasp_wid_mask=None):
        if self.training:
            raise NotImplementedError("crog_b200.CROG implements the inference path only; call .eval()")
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != img.shape[3]:
            raise RuntimeError(f"expected img of shape [B,3,S,S], got {tuple(img.shape)}")
        if word.dim() != 2 or word.shape[0] != img.shape[0] or word.shape[1] != self.cfg.word_len:
            raise RuntimeError(f"expected word of shape [{img.shape[0]},{self.cfg.word_len}], got {tuple(word.shape)}")
        B, S = img.shape[0], img.shape[-1]
        plan = self.plan_for(B, S)
        with torch.cuda.device(plan.dev):
            plan.img.copy_(img, non_blocking=True)
            plan.word.copy_(word, non_blocking=True)
            self._run(plan)
            out = plan.out.clone()
        maps = tuple(out[i] for i in range(plan.NH))
        if self.use_grasp_masks:
            return maps, (mask, grasp_qua_mask, grasp_sin_mask, grasp_cos_mask, grasp_wid_mask)
        return maps[0], mask

    def _run(self, plan):
        if not self.use_cuda_graph:
            plan.run()
            return
        key = id(plan)
        g = self._graphs.get(key)
        if g is None:
            plan.run()  # eager warm-up: sets kernel attributes, faults surface here with op context
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                plan.run()
            self._graphs[key] = g
        g.replay()
