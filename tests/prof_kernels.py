"""Profiling helper (not a test): one launch of each memory-bound / tail / SSG kernel at the benchmark's shapes, bracketed
by cudaProfilerStart/Stop.  usage: ncu --set full --profile-from-start off -k regex:<kernel> -c 1 ... python tests/prof_kernels.py <which>
which: glue | lnchain | gaussian | warp | preprocess | ssgpost | tail_blobs | tail_stress"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crog_b200 import _lib as L  # noqa: E402
from crog_b200 import synth  # noqa: E402
from crog_b200.engine import postprocess  # noqa: E402
from crog_b200.utils import grasp_eval as GE  # noqa: E402
from crog_b200.utils import warp as WP  # noqa: E402

which = sys.argv[1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(0)
B = 64


def bracket(fn, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if which == "glue":  # sigmoid + bicubic x4 of the five 104x104 logit maps of 64 samples
    maps = torch.randn((5, B, 104, 104), generator=g, device=dev)
    bracket(lambda: postprocess(maps, (416, 416)))
elif which == "lnchain":  # decoder: vis += LN(x); v2 = LN(vis) on 64 x 676 rows of 512
    rows, D = B * 676, 512
    x = torch.randn((rows, D), generator=g, device=dev).to(torch.bfloat16)
    vis = torch.randn((rows, D), generator=g, device=dev)
    z = torch.empty((rows, D), device=dev, dtype=torch.bfloat16)
    w = [torch.rand(D, device=dev) for _ in range(4)]
    lib = L.lib()
    bracket(lambda: L.check(lib.crog_layernorm_chain(x.data_ptr(), L.BF16, w[0].data_ptr(), w[1].data_ptr(), vis.data_ptr(), vis.data_ptr(),
                                                     w[2].data_ptr(), w[3].data_ptr(), z.data_ptr(), L.BF16, rows, D, 1e-5, L.stream_ptr())))
elif which == "gaussian":  # 768 instance quality maps 480x640 (SSG batch 64 x 12 instances)
    maps = torch.rand((768, 480, 640), generator=g, device=dev)
    bracket(lambda: GE.gaussian_batched(maps, 2.0))
elif which == "warp":  # inverse letterbox of the five post-processed maps of 64 samples to 480x640
    post = torch.rand((5, B, 416, 416), generator=g, device=dev)
    mat, mat_inv = WP.get_transform_mat((480, 640), (416, 416), inverse=True)
    aff = WP.device_affine(mat_inv, B, dev)
    bracket(lambda: WP.warp_affine_cubic(post, aff, (640, 480), 0.0))
elif which == "preprocess":  # uint8 480x640 frames -> normalised 416x416 network input
    frames = synth.make_frames_u8(B).to(dev)
    mat, _ = WP.get_transform_mat((480, 640), (416, 416), inverse=True)
    aff = WP.device_affine(mat, B, dev)
    bracket(lambda: WP.preprocess_images(frames, aff, (416, 416)))
elif which == "ssgpost":
    cfg = synth.ssg_cfg()
    ods = [synth.make_ssg_output_dict(cfg, n_confident=8, seed=100 + i) for i in range(16)]
    od = {k: torch.cat([o[k] for o in ods]).to(dev) for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
    od["anchors"] = ods[0]["anchors"]
    bracket(lambda: GE.ssg_post_processing_batched(cfg, od, (480, 640)))
elif which in ("tail_blobs", "tail_stress"):
    n = 4096
    gt, cnt = synth.make_gt_rects(n, 64, seed=4)
    d_gt, d_cnt = torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev)
    kind = which.split("_")[1]
    q, s, c, w = bench.gen_tail_maps_device(n, kind, 7 if kind == "blobs" else 8, dev)

    def run():
        pk, npk, gr = GE.detect_grasps_batched(q, s, c, w, 5)
        GE.jacquard_batched(gr, npk, d_gt, d_cnt)

    bracket(run)
else:
    raise SystemExit(f"unknown kernel set {which}")
print("ok", which)
