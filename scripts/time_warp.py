"""Event-timed inverse letterbox of the five post-processed maps of 64 samples (416x416 -> 480x640), v2 vs v1 (CROG_WARP_V1)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200.utils import warp as WP
dev = torch.device("cuda", 0)
B = 64
post = torch.rand((5, B, 416, 416), device=dev)
mat, mat_inv = WP.get_transform_mat((480, 640), (416, 416), inverse=True)
aff = WP.device_affine(mat_inv, B, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t():
    ts = []
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = WP.warp_affine_cubic(post, aff, (640, 480), 0.0); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2], out
t2, o2 = t()
os.environ["CROG_WARP_V1"] = "1"
t1, o1 = t()
print("warp us: v2 %.1f  v1 %.1f  bit-identical %s" % (t2, t1, bool(torch.equal(o1, o2))))
