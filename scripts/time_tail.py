"""Event-timed crog_detect_grasps on 1024 maps 416x416 of one kind: iid uniform / quantised to 1/16 / blobs (K = 5)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.utils import grasp_eval as GE

dev = torch.device("cuda", 0)
n = 1024
g = torch.Generator(device=dev).manual_seed(0)
base = torch.rand((n, 416, 416), generator=g, device=dev)
kinds = {"iid": base, "quant16": torch.floor(base * 16) / 16, "quant256": torch.floor(base * 256) / 256}
qb = synth.make_tail_maps(64, "blobs", seed=3)[0]
if True:
    kinds["blobs"] = torch.from_numpy(qb).to(dev).repeat(n // 64, 1, 1)
s = torch.randn((n, 416, 416), generator=g, device=dev); c = torch.randn((n, 416, 416), generator=g, device=dev)
w = torch.rand((n, 416, 416), generator=g, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, q in kinds.items():
    if only and name != only: continue
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); GE.detect_grasps_batched(q, s, c, w, 5); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{name}: {ts[len(ts)//2]:.1f} us per {n} maps ({n*416*416*4/ts[len(ts)//2]/1e3:.0f} GB/s of q)")
