"""Event-timed Gaussian (sigma 2) over 768 maps 480x640: one-kernel form vs the two-pass form (in place)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.utils import grasp_eval as GE
dev = torch.device("cuda", 0)
maps = torch.rand((768, 480, 640), device=dev)
out = torch.empty_like(maps)
def t(fn, n=8):
    ts = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
fused = t(lambda: GE.gaussian_batched(maps, 2.0, out=out))
ref = out.clone()
work = maps.clone()
two = t(lambda: GE.gaussian_batched(work, 2.0, out=work))
work.copy_(maps); GE.gaussian_batched(work, 2.0, out=work); torch.cuda.synchronize()
print("fused %.3f ms   two-pass (in place, incl. tmp alloc) %.3f ms   bit-identical %s" % (fused, two, bool(torch.equal(ref, work))))
