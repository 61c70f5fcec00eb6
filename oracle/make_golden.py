"""ORACLE tooling: pin oracle/crog_forward.py against the real reference and write
tests/golden/model_*.npz.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py

What it does: imports the *unmodified* reference modules from
/root/reference, stubs ``torch.jit.load`` (the CLIP RN50.pt archive is not available
offline) with a seeded ``CLIP(...)`` so ``CROG(cfg)`` constructs, loads our synthetic
state-dict with ``strict=True`` (which also pins the name/shape table of
crog_b200/spec.py), runs ``model(img, word)`` in eval mode on CPU fp32, asserts the
restatement in oracle/crog_forward.py matches, and stores inputs-by-seed + outputs.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from crog_b200 import synth  # noqa: E402
from oracle import crog_forward as O  # noqa: E402


def build_reference(cfg):
    sys.path.insert(0, REF)
    import model.clip as rclip  # reference
    from model.crog import CROG as RefCROG  # reference

    class _Stub:
        def __init__(self, sd):
            self._sd = sd

        def eval(self):
            return self

        def state_dict(self):
            return self._sd

    real_load = torch.jit.load

    def fake_load(path, map_location=None):
        torch.manual_seed(0)
        m = rclip.CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, cfg.word_len, 49408, 512, 8, 12)
        return _Stub(m.state_dict())

    torch.jit.load = fake_load
    try:
        ref = RefCROG(cfg).eval()
    finally:
        torch.jit.load = real_load
        sys.path.remove(REF)
    return ref


def run_case(tag, word_len, batch, mode, seed_w, size=416, **cfg_over):
    """cfg_over: the ablation switches of config/OCID-VLG/crog_multiple_r50_wo_contrastive.yaml (use_contrastive=False:
    no TransformerDecoder, model/crog.py:27-39,70-72) and ..._wo_grasps.yaml (use_grasp_masks=False: the mask-only
    ``Projector`` of model/layers.py:135-173, eval return ``(pred, mask)``, model/crog.py:115-133)."""
    cfg = synth.default_cfg(word_len=word_len, **cfg_over)
    ref = build_reference(cfg)
    sd = synth.make_state_dict(cfg, seed=seed_w, mode=mode)
    ref_sd = ref.state_dict()
    assert set(ref_sd) == set(sd), (set(ref_sd) ^ set(sd))
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    ref.load_state_dict(sd, strict=True)
    img, word = synth.make_inputs(batch, word_len, size=size)  # size != 416: the attention pool resizes its positional embedding (clip.py:101-104)
    t0 = time.time()
    with torch.no_grad():
        (r_maps, _) = ref(img, word)
        if torch.is_tensor(r_maps):  # mask-only ablation returns the bare prediction
            r_maps = (r_maps,)
        # intermediates straight from the reference sub-modules
        c3, c4, c5 = ref.backbone.encode_image(img)
        wfeat, state = ref.backbone.encode_text(word)
        fq = ref.neck((c3, c4, c5), state)
    t_ref = time.time() - t0
    o_maps, inter = O.crog_forward(sd, cfg, img, word, keep=True)
    errs = {}
    for nm, a, b in [("c3", c3, inter["c3"]), ("c4", c4, inter["c4"]), ("c5", c5, inter["c5"]),
                     ("word", wfeat, inter["word"]), ("state", state, inter["state"]),
                     ("fq_neck", fq, inter["fq_neck"])]:
        errs[nm] = float((a - b).abs().max())
    assert len(r_maps) == len(o_maps) == (5 if cfg.use_grasp_masks else 1)
    for i, nm in enumerate(("mask", "qua", "sin", "cos", "wid")[:len(r_maps)]):
        errs[nm] = float((r_maps[i] - o_maps[i]).abs().max())
    print(tag, "ref fwd %.2fs" % t_ref, {k: "%.2e" % v for k, v in errs.items()})
    # the restatement must agree with the reference to fp32 round-off: 1e-4 abs on logits of range +-16; the decoder-less
    # ablation feeds the projector un-normalised features (logits up to +-104), so the bar scales with the range
    rng_scale = max(1.0, max(float(m.abs().max()) for m in r_maps) / 16.0)
    assert all(v <= 1e-4 * rng_scale for v in errs.values()), (errs, rng_scale)
    post = O.postprocess(r_maps, (size, size))
    out = {
        "word_len": np.int64(word_len), "batch": np.int64(batch), "seed_w": np.int64(seed_w), "size": np.int64(size),
        "maps": torch.stack([m[:, 0] for m in r_maps], 1).numpy().astype(np.float32),  # B,5,104,104
        "state": state.numpy(), "word_feat": wfeat.numpy(),
        "c5_sample": c5[:, ::16].numpy(), "c4_sample": c4[:, ::64].numpy(), "c3_sample": c3[:, ::64, ::2, ::2].numpy(),
        "fq_neck_sample": fq[:, ::16].numpy(),
    }
    if cfg.use_contrastive:
        out["fq_dec_sample"] = inter["fq_dec"][:, ::16].numpy()
    if cfg.use_grasp_masks:
        out["post_qua_rowsum"] = post[1].sum(-1).numpy()
    out["use_contrastive"], out["use_grasp_masks"] = np.bool_(cfg.use_contrastive), np.bool_(cfg.use_grasp_masks)
    path = os.path.join(ROOT, "tests", "golden", f"model_{tag}.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, mode=np.array(mode), **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    run_case("L17_perturbed", 17, 2, "perturbed", 0)
    run_case("L20_init", 20, 1, "init", 0)
    run_case("L17_wo_contrastive", 17, 1, "perturbed", 0, use_contrastive=False)
    run_case("L17_wo_grasps", 17, 1, "perturbed", 0, use_grasp_masks=False)
    run_case("L17_perturbed_s320", 17, 1, "perturbed", 0, size=320)
