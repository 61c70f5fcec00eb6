"""Event-timed crog_preprocess_u8: 64 uint8 480x640 frames -> normalised 416x416 network input."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200.utils import warp as WP
dev = torch.device("cuda", 0)
B = 64
frames = synth.make_frames_u8(B).to(dev)
mat, _ = WP.get_transform_mat((480, 640), (416, 416), inverse=True)
aff = WP.device_affine(mat, B, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(14):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = WP.preprocess_images(frames, aff, (416, 416)); b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print("preprocess us median %.1f min %.1f  checksum %.6f" % (ts[len(ts) // 2], ts[0], float(out.double().sum())))
