"""Event-timed glue launch (sigmoid + bicubic x4 of the five 104x104 maps of 64 samples); L2 flushed between launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.engine import postprocess

dev = torch.device("cuda", 0)
maps = torch.randn((5, 64, 104, 104), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(25):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); postprocess(maps, (416, 416)); b.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print("glue us median %.1f min %.1f" % (ts[len(ts) // 2], ts[0]))
