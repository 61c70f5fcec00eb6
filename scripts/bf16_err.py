"""Per-map bf16 error of the tcgen05 forward against the reference goldens (prints; tests assert the bounds)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.model import CROG
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_L17_perturbed.npz"))
cfg = synth.default_cfg(17)
model = CROG(cfg, precision="bf16"); model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed")); model = model.cuda()
img, word = synth.make_inputs(2, 17)
maps, _ = model(img.cuda(), word.cuda()); torch.cuda.synchronize()
got = torch.stack([m[:, 0] for m in maps], 1).cpu().numpy(); ref = g["maps"]
print("FOLD", os.environ.get("CROG_FFN_LN_FOLD", "1"), "rel-L2 per map", [round(float(np.linalg.norm(got[:, i] - ref[:, i]) / np.linalg.norm(ref[:, i])), 4) for i in range(5)],
      "max-abs", round(float(np.abs(got - ref).max()), 4))
