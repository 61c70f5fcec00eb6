"""Per-op CUDA-event table of the SSG forward plan (eager replay), batch 64."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crog_b200 import synth
from crog_b200.model import SSG
dev = torch.device("cuda", 0)
cfg = synth.ssg_cfg()
model = SSG(cfg, precision="bf16"); model.load_state_dict(synth.make_ssg_state_dict(cfg, 0, "perturbed"), strict=True); model = model.to(dev)
B = 64
rgb, depth = synth.make_ssg_inputs(B, cfg.img_size)
model({"rgb": rgb.to(dev), "depth": depth.to(dev)}); torch.cuda.synchronize()
plan = model.plan_for(B)
durs = bench._op_durations(plan, reps=3) * 1e3  # ms -> us
rows = sorted(zip(durs, plan.op_names), reverse=True)
tot = float(durs.sum())
print("total %.2f ms, %d ops" % (tot / 1e3, len(durs)))
for d, n in rows[:40]:
    fl = plan.gemm_alg_flops.get(n)
    print("%8.1f us %5.1f%%  %-42s %s" % (d, 100 * d / tot, n[:42], ("%.0f TF/s" % (fl / d / 1e6)) if fl else ""))
import collections
agg = collections.defaultdict(float)
for d, n in zip(durs, plan.op_names): agg[n.split(".")[0]] += d
print({k: round(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
