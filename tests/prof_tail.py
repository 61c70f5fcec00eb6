"""Profiling helper (not a test): one pass of the config-5 tail (detect + Jaccard) per distribution.
usage: ncu ... python tests/prof_tail.py [n_maps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crog_b200 import synth  # noqa: E402
from crog_b200.utils import grasp_eval as GE  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
gt, cnt = synth.make_gt_rects(n, 64, seed=4)
d_gt, d_cnt = torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev)
for kind, seed in (("blobs", 7), ("stress", 8)):
    q, s, c, w = bench.gen_tail_maps_device(n, kind, seed, dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    peaks, npk, grasps = GE.detect_grasps_batched(q, s, c, w, 5)
    GE.jacquard_batched(grasps, npk, d_gt, d_cnt, counters=counters)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(kind, counters.tolist())
    del q, s, c, w
