"""Golden checksums for the CLIP-archive ingest row, produced by the UNMODIFIED reference
``model.clip.build_model(state_dict, txt_length, load_weights=True).float()`` imported from /root/reference.

Input: the ``backbone.*`` part of the seeded synthetic state-dict (crog_b200.synth, "perturbed", seed 0) with the prefix
stripped - fp32 values that are NOT representable in fp16, so the round trip is visible.  Output: for every tensor of
the reference's resulting state-dict, (sum, sum of |x|) in float64, written as hex floats to
tests/golden/clip_ingest_checksums.json.

    python oracle/make_golden_ingest.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crog_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "clip_ingest_checksums.json")


def checksum(t: torch.Tensor):
    d = t.detach().double().flatten()
    return [float(d.sum()).hex(), float(d.abs().sum()).hex()]


def main():
    sys.path.insert(0, "/root/reference")
    from model.clip import build_model

    cfg = synth.default_cfg(17)
    sd = synth.make_state_dict(cfg, 0, "perturbed")
    clip_sd = {k[len("backbone."):]: v.clone() for k, v in sd.items() if k.startswith("backbone.") and "attnpool.connect" not in k}
    clip_sd["input_resolution"] = torch.tensor(224)
    clip_sd["context_length"] = torch.tensor(77)
    clip_sd["vocab_size"] = torch.tensor(49408)
    torch.manual_seed(0)
    model = build_model(dict(clip_sd), 17, True).float()
    ref = model.state_dict()
    out = {k: checksum(v) for k, v in ref.items() if "attnpool.connect" not in k and not k.endswith("num_batches_tracked")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", OUT, len(out), "tensors")


if __name__ == "__main__":
    main()
