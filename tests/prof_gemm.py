"""Profiling helper (not a test): one large 3x3 convolution GEMM (proj.vis.3 shape at batch 8) per configuration.
usage: ncu ... python tests/prof_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import conv_w, pad_nhwc, run_gemm  # noqa: E402

torch.manual_seed(0)
B, H, W, Cin, Cout = 8, 104, 104, 512, 256
dt = torch.bfloat16
x = torch.randn(B, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (Cin * 9) ** -0.5
a = pad_nhwc(x, dt)
wk = conv_w(w, dt)
out = torch.zeros((B * (H + 2) * (W + 2), Cout), device="cuda", dtype=dt)


def run():
    run_gemm(a, wk, Cout, out, taps=9, H=H, W=W, in_padded=True, out_padded=True, sample_rows=(H + 2) * (W + 2), act=L.ACT_RELU,
             impl=L.IMPL_TCGEN05)


for env in ({}, {"CROG_GEMM_PAIR": "1"}):
    for k in ("CROG_GEMM_PAIR",):
        os.environ.pop(k, None)
    os.environ.update(env)
    run(); run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / 5
    gf = 2 * B * H * W * Cout * 9 * Cin / 1e9
    print(env, f"{ms:.4f} ms  {gf / ms:.1f} TF/s (incl. host launch + sync gaps)")
