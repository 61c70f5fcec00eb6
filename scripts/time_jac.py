"""Event-timed crog_jaccard on 4096 samples (5 predictions from 'blobs' maps x 64 GT rectangles)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crog_b200 import synth
from crog_b200.utils import grasp_eval as GE
dev = torch.device("cuda", 0)
n = 4096
gt, cnt = synth.make_gt_rects(n, 64, seed=4)
d_gt, d_cnt = torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev)
q, s, c, w = bench.gen_tail_maps_device(n, "blobs", 7, dev)
_, npk, grasps = GE.detect_grasps_batched(q, s, c, w, 5)
ts = []
for i in range(12):
    g2 = d_gt.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); flags = GE.jacquard_batched(grasps, npk, g2, d_cnt); b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
fl = flags[0] if isinstance(flags, (tuple, list)) else flags
print("jaccard us median %.1f min %.1f  checksum %d" % (ts[len(ts) // 2], ts[0], int(torch.as_tensor(fl).sum().item())))
