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
import threading
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from ..spec import crog_tensor_specs
from .. import _lib as L


_capture_streams: Dict[int, torch.cuda.Stream] = {}


def _capture_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _capture_streams.get(dev.index)
    if s is None:
        s = _capture_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


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
        # Plans own their activation buffers (about 95 MB per sample at 416x416): the cache is a small LRU, a batch smaller
        # than a cached plan's runs on that plan (samples are independent end to end), and `prepare()` moves plan build +
        # tile autotune + graph capture out of the first forward.  The dicts are shared by DataParallel replicas (shallow
        # __dict__ copies), so they are keyed by device and guarded by one lock.
        self.max_plans = int(os.environ.get("CROG_MAX_PLANS", "2"))
        self._plans: "OrderedDict[Tuple, object]" = OrderedDict()
        self._graphs: Dict[Tuple, object] = {}
        self._lock = threading.RLock()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())
        self.eval()

    # ------------------------------------------------------------------ weight / plan management
    def invalidate(self):
        """Drop packed weights and captured graphs (call after changing parameters in place)."""
        with self._lock:
            self._plans.clear()
            self._graphs.clear()

    def __getstate__(self):
        # plans / graphs / the lock are per-process run-time state: a pickled or deep-copied module starts without them
        d = dict(self.__dict__)
        d["_plans"], d["_graphs"], d["_lock"] = OrderedDict(), {}, None
        return d

    def __setstate__(self, state):
        super().__setstate__(state)
        self._lock = threading.RLock()

    def _full_state(self) -> Dict[str, torch.Tensor]:
        """``state_dict()`` that also works on a ``torch.nn.DataParallel`` replica, whose ``_parameters`` are empty (the
        per-device copies are plain attributes, torch/nn/parallel/replicate.py): walk the name table with getattr."""
        out = {}
        for s_ in crog_tensor_specs(self.cfg):
            m = self
            parts = s_.name.split(".")
            for p_ in parts[:-1]:
                m = m._modules[p_]
            t = m._parameters.get(parts[-1])
            if t is None:
                t = m._buffers.get(parts[-1])
            if t is None:
                t = getattr(m, parts[-1])
            out[s_.name] = t.detach()
        return out

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

    def _device(self) -> torch.device:
        m = self._modules["backbone"]
        t = m._parameters.get("text_projection")
        return (t if t is not None else getattr(m, "text_projection")).device

    def plan_for(self, batch: int, size: int, keep: bool = False, exact: bool = True):
        """The execution plan serving ``batch`` samples of ``size`` x ``size``.  ``exact=False`` accepts a cached plan of a
        larger batch (the smallest one): a dataloader's ragged last batch then costs neither 95 MB/sample of new buffers
        nor a tile autotune; it runs in the first rows of the big plan."""
        from .plan import ForwardPlan

        dev = self._device()
        if dev.type != "cuda":
            raise L.CrogError("crog_b200.CROG.forward needs the module on a CUDA device (sm_100a); no CPU fallback exists")
        with self._lock:
            key = (batch, size, self.precision, self.gemm_impl, keep, dev.index)
            plan = self._plans.get(key)
            if plan is None and not exact:
                fits = [k for k in self._plans if k[1:] == key[1:] and k[0] > batch]
                if fits:
                    key = min(fits)
                    plan = self._plans[key]
            if plan is None:
                with torch.cuda.device(dev):
                    plan = ForwardPlan(self._full_state(), self.cfg, batch, self.precision, dev, size, self.gemm_impl, keep)
                    if self.autotune and os.environ.get("CROG_AUTOTUNE", "1") != "0":
                        plan.autotune()
                self._plans[key] = plan
                per_dev = [k for k in self._plans if k[-1] == dev.index]
                while len(per_dev) > max(self.max_plans, 1):  # least recently used plan of this device goes, with its graph
                    old = per_dev.pop(0)
                    self._graphs.pop(id(self._plans.pop(old)), None)
            self._plans.move_to_end(key)
        return plan

    def prepare(self, batch: int, size: Optional[int] = None):
        """Build the plan for ``batch`` (weight packing, buffer allocation, per-layer tile autotune: about 2 s at batch 64)
        and capture its CUDA graph now, instead of inside the first ``forward``.  Call once with the largest batch the
        loader yields; smaller batches reuse that plan."""
        plan = self.plan_for(batch, size or self.cfg.input_size)
        with torch.cuda.device(plan.dev):
            self._run(plan)
            torch.cuda.synchronize()
        return self

    def input_buffer(self, batch: int, size: Optional[int] = None) -> torch.Tensor:
        """The float32 [B,3,S,S] tensor the plan for this batch reads.  A producer on the same stream (e.g.
        ``utils.warp.preprocess_images(..., out=model.input_buffer(B))``) can fill it in place and pass it to ``forward``,
        which then skips its 2 MB-per-sample device-to-device input copy."""
        return self.plan_for(batch, size or self.cfg.input_size, exact=False).img[:batch]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, img, word, mask=None, grasp_qua_mask=None, grasp_sin_mask=None, grasp_cos_mask=None,
                grasp_wid_mask=None):
        out = self.forward_stacked(img, word)
        maps = tuple(out[i] for i in range(out.shape[0]))
        if self.use_grasp_masks:
            return maps, (mask, grasp_qua_mask, grasp_sin_mask, grasp_cos_mask, grasp_wid_mask)
        return maps[0], mask

    @torch.no_grad()
    def forward_stacked(self, img, word) -> torch.Tensor:
        """The forward's logits as ONE tensor [NH, B, 1, S/4, S/4] (NH = 5: mask, qua, sin, cos, wid; 1 for the mask-only
        model) - what ``forward`` returns as a tuple of views.  The evaluator takes this form so that the glue kernel reads
        the heads in place instead of re-stacking them."""
        if self.training:
            raise NotImplementedError("crog_b200.CROG implements the inference path only; call .eval()")
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != img.shape[3]:
            raise RuntimeError(f"expected img of shape [B,3,S,S], got {tuple(img.shape)}")
        if word.dim() != 2 or word.shape[0] != img.shape[0] or word.shape[1] != self.cfg.word_len:
            raise RuntimeError(f"expected word of shape [{img.shape[0]},{self.cfg.word_len}], got {tuple(word.shape)}")
        B, S = img.shape[0], img.shape[-1]
        plan = self.plan_for(B, S, exact=False)
        with torch.cuda.device(plan.dev):
            if img.data_ptr() != plan.img.data_ptr():  # input_buffer(): the producer wrote the plan's input in place
                plan.img[:B].copy_(img, non_blocking=True)
            plan.word[:B].copy_(word, non_blocking=True)  # rows >= B keep an earlier batch: independent samples, results unused
            self._run(plan)
            return plan.out[:, :B].clone()

    def _run(self, plan):
        if not self.use_cuda_graph:
            plan.run()
            return
        key = id(plan)
        g = self._graphs.get(key)
        if g is None:
            with self._lock:  # one capture at a time (DataParallel replicas run forward from worker threads)
                g = self._graphs.get(key)
                if g is None:
                    plan.run()  # eager warm-up: sets kernel attributes, faults surface here with op context
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    # torch.cuda.graph's default capture stream is ONE process-wide stream on whichever device captured
                    # first; entering it from another device's replica would switch the current device: use our own
                    with torch.cuda.graph(g, stream=_capture_stream(plan.dev), capture_error_mode="thread_local"):
                        plan.run()
                    self._graphs[key] = g
        g.replay()
