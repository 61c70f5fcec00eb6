"""Profiling helper (not a test): one eager forward of the bench workload, launches in plan order.
usage: CROG_NO_FORK=1 ncu ... python tests/prof_forward.py [B] [op-names-out.tsv]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CROG_NO_FORK", "1")
from crog_b200 import synth  # noqa: E402
from crog_b200.model import CROG  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = synth.default_cfg(17)
model = CROG(cfg, precision="bf16", use_cuda_graph=False)
model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed"))
model = model.cuda()
img, word = synth.make_inputs(B, 17)
plan = model.plan_for(B, 416)
plan.img.copy_(img.cuda()); plan.word.copy_(word.cuda())
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", plan.n_launches)
if len(sys.argv) > 2:  # op names, one line per kernel launch in plan order (profiles/launch_list.py joins them with the ncu rows)
    with open(sys.argv[2], "w") as f:
        for name, n in zip(plan.op_names, plan.op_launches):
            tc = plan.tile_choice.get(name)
            for _ in range(n):
                f.write(f"{name}\t{tc[0] if tc else ''}\n")
