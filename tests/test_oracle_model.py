"""Pins oracle/crog_forward.py (the CPU restatement of the reference forward) against
the golden vectors generated from the real reference by oracle/make_golden.py."""
import os

import numpy as np
import pytest
import torch

from crog_b200 import synth
from crog_b200.spec import crog_tensor_specs
from oracle import crog_forward as O


def test_spec_table_counts():
    specs = crog_tensor_specs(synth.default_cfg(17))
    assert len(specs) == 662  # 449 parameters + 213 buffers
    assert sum(1 for s in specs if s.is_buffer) == 213
    assert len({s.name for s in specs}) == 662


@pytest.mark.parametrize("tag", ["L20_init", "L17_perturbed", "L17_perturbed_s320"])
def test_restatement_matches_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"model_{tag}.npz"))
    L, B = int(g["word_len"]), int(g["batch"])
    S = int(g["size"]) if "size" in g.files else 416  # 320: the attention pool's resized positional embedding
    cfg = synth.default_cfg(L)
    sd = synth.make_state_dict(cfg, int(g["seed_w"]), str(g["mode"]))
    img, word = synth.make_inputs(B, L, size=S)
    torch.set_num_threads(os.cpu_count())
    maps, inter = O.crog_forward(sd, cfg, img, word, keep=True)
    got = torch.stack([m[:, 0] for m in maps], 1).numpy()
    assert np.abs(got - g["maps"]).max() <= 1e-4
    assert np.abs(inter["state"].numpy() - g["state"]).max() <= 1e-4
    assert np.abs(inter["word"].numpy() - g["word_feat"]).max() <= 1e-4
    assert np.abs(inter["c5"][:, ::16].numpy() - g["c5_sample"]).max() <= 1e-4
    assert np.abs(inter["fq_dec"][:, ::16].numpy() - g["fq_dec_sample"]).max() <= 1e-4
    post = O.postprocess(maps, (S, S))
    assert np.abs(post[1].sum(-1).numpy() - g["post_qua_rowsum"]).max() <= 1e-2


@pytest.mark.parametrize("tag,over", [("L17_wo_contrastive", {"use_contrastive": False}), ("L17_wo_grasps", {"use_grasp_masks": False})])
def test_restatement_matches_reference_golden_ablations(golden_dir, tag, over):
    """config/OCID-VLG/crog_multiple_r50_wo_contrastive.yaml / ..._wo_grasps.yaml shapes (model/crog.py:25-45)."""
    g = np.load(os.path.join(golden_dir, f"model_{tag}.npz"))
    L, B = int(g["word_len"]), int(g["batch"])
    cfg = synth.default_cfg(L, **over)
    sd = synth.make_state_dict(cfg, int(g["seed_w"]), str(g["mode"]))
    assert not over.get("use_contrastive", True) == (not any(k.startswith("decoder") for k in sd))
    img, word = synth.make_inputs(B, L)
    torch.set_num_threads(os.cpu_count())
    maps, inter = O.crog_forward(sd, cfg, img, word, keep=True)
    got = torch.stack([m[:, 0] for m in maps], 1).numpy()
    assert got.shape == g["maps"].shape and got.shape[1] == (5 if cfg.use_grasp_masks else 1)
    scale = max(1.0, float(np.abs(g["maps"]).max()) / 16.0)  # the decoder-less logits reach +-104
    assert np.abs(got - g["maps"]).max() <= 1e-4 * scale
    assert np.abs(inter["fq_neck"][:, ::16].numpy() - g["fq_neck_sample"]).max() <= 1e-4
