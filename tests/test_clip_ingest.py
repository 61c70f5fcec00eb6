"""CLIP-archive ingest row: crog_b200.model.clip_ingest against checksums of the state-dict the
unmodified reference ``build_model(sd, txt_length, load_weights=True).float()`` produces from the same seeded input
(oracle/make_golden_ingest.py).  Bit-exact: the fp16 round trip is deterministic."""
import json
import os

import pytest
import torch

from crog_b200 import synth
from crog_b200.model.clip_ingest import ARCHIVE_ONLY_KEYS, FP16_ROLES, clip_state_to_backbone
from crog_b200.spec import crog_tensor_specs

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "clip_ingest_checksums.json")))


def _clip_sd(cfg):
    sd = synth.make_state_dict(cfg, 0, "perturbed")
    clip = {k[len("backbone."):]: v.clone() for k, v in sd.items() if k.startswith("backbone.") and "attnpool.connect" not in k}
    for k in ARCHIVE_ONLY_KEYS:
        clip[k] = torch.tensor(1)
    return sd, clip


def test_ingest_matches_reference_build_model():
    cfg = synth.default_cfg(17)
    sd, clip = _clip_sd(cfg)
    out, missing, unexpected = clip_state_to_backbone(clip, cfg)
    assert unexpected == []
    assert all("attnpool.connect" in m for m in missing) and len(missing) == 6
    checked = 0
    for k, (s_hex, a_hex) in GOLD.items():
        t = out["backbone." + k].double().flatten()
        assert float(t.sum()) == float.fromhex(s_hex) and float(t.abs().sum()) == float.fromhex(a_hex), k
        checked += 1
    assert checked == len(GOLD) >= 430
    # the round trip is visible on this input: converted roles changed, the others did not
    roles = {s.name: s.role for s in crog_tensor_specs(cfg)}
    for name, t in out.items():
        if t.dtype != torch.float32:
            continue
        same = torch.equal(t, sd[name].float())
        assert same == (roles[name] not in FP16_ROLES) or t.numel() == 1, name


def test_ingest_rejects_wrong_shapes_and_reports_unknown_keys():
    cfg = synth.default_cfg(17)
    _, clip = _clip_sd(cfg)
    clip["visual.proj"] = torch.zeros(3)
    out, _, unexpected = clip_state_to_backbone(clip, cfg)
    assert unexpected == ["visual.proj"]
    clip["visual.conv1.weight"] = torch.zeros(8, 3, 3, 3)
    with pytest.raises(RuntimeError):
        clip_state_to_backbone(clip, cfg)


def test_load_clip_updates_module_parameters():
    from crog_b200.model import CROG

    cfg = synth.default_cfg(17)
    _, clip = _clip_sd(cfg)
    m = CROG(cfg)
    before = m.state_dict()["backbone.visual.attnpool.connect.0.weight"].clone()
    missing, unexpected = m.load_clip(clip)
    sd = m.state_dict()
    want, _, _ = clip_state_to_backbone(clip, cfg)
    assert torch.equal(sd["backbone.visual.conv1.weight"], want["backbone.visual.conv1.weight"])
    assert torch.equal(sd["backbone.visual.attnpool.connect.0.weight"], before)  # absent from the archive: keeps its init
    assert len(missing) == 6 and unexpected == []
