"""Ingest of a CLIP RN50 archive into the drop-in module.

The reference builds its backbone as ``build_model(torch.jit.load(cfg.clip_pretrain).state_dict(), cfg.word_len,
cfg.use_pretrained_clip).float()`` (model/crog.py:20-23, model/clip.py:503-556).  With ``use_pretrained_clip`` that is:
construct CLIP with the hyper-parameters read off the archive, ``convert_weights`` (every Conv / Linear /
MultiheadAttention weight and bias and ``text_projection`` become fp16, model/clip.py:477-500), copy the archive in
with ``strict=False`` (the copy casts to the destination dtype, so those tensors pass through fp16 whatever the archive
stores; BatchNorm / LayerNorm / embeddings / positional embeddings keep the archive's value), then ``.float()``.
Restated here on plain tensors: one rounding through fp16 for the roles ``convert_weights`` touches, a plain fp32 cast
for the rest.  ``attnpool.connect`` is CROG's own addition, absent from the archive: it keeps the module's init, as in
the reference.  Trained CROG checkpoints (fp32 ``.pth``, ``module.`` prefix or not) go through
``CROG.load_state_dict`` directly and need none of this.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ..spec import crog_tensor_specs

# roles (crog_b200/spec.py) that model/clip.py:convert_weights stores in fp16
FP16_ROLES = ("conv", "linear_w", "bias", "proj")
ARCHIVE_ONLY_KEYS = ("input_resolution", "context_length", "vocab_size")  # dropped by build_model (clip.py:547-549)


def clip_state_to_backbone(clip_sd: Dict[str, torch.Tensor], cfg) -> Tuple[Dict[str, torch.Tensor], list, list]:
    """CLIP state-dict (archive names: ``visual.*``, ``transformer.*``, ``token_embedding.weight`` ...) ->
    ({"backbone.<name>": fp32 tensor}, missing backbone names, unexpected archive names), with the reference's dtype
    round trip.  Shapes are checked against the module's table; a mismatch raises like load_state_dict does."""
    specs = {s.name: s for s in crog_tensor_specs(cfg) if s.name.startswith("backbone.")}
    out, unexpected = {}, []
    for k, v in clip_sd.items():
        if k in ARCHIVE_ONLY_KEYS:
            continue
        s = specs.get("backbone." + k)
        if s is None:
            unexpected.append(k)
            continue
        v = v.detach().to("cpu")
        if tuple(v.shape) != tuple(s.shape):
            raise RuntimeError(f"size mismatch for backbone.{k}: archive {tuple(v.shape)} vs module {tuple(s.shape)}")
        if s.dtype == "int64":
            out[s.name] = v.to(torch.int64)
        elif s.role in FP16_ROLES:
            out[s.name] = v.to(torch.float16).to(torch.float32)
        else:
            out[s.name] = v.to(torch.float32)
    missing = [n for n in specs if n not in out]
    return out, missing, unexpected


def load_clip_archive(path: str) -> Dict[str, torch.Tensor]:
    """State-dict of a CLIP TorchScript archive (``RN50.pt``), as the reference reads it (model/crog.py:20-21)."""
    return torch.jit.load(path, map_location="cpu").eval().state_dict()
