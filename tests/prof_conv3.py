"""Profiling helper (not a test): one 3x3 convolution on the zero-haloed layout at a forward shape.
usage: [ncu ...] python tests/prof_conv3.py B H W Cin Cout [tile_cfg] [out_padded=1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import run_gemm  # noqa: E402

B, H, W, Cin, Cout = [int(v) for v in sys.argv[1:6]]
cfg = int(sys.argv[6]) if len(sys.argv) > 6 else 0
out_padded = int(sys.argv[7]) if len(sys.argv) > 7 else 1
torch.manual_seed(0)
dt = torch.bfloat16
rows = B * (H + 2) * (W + 2)
a = (torch.randn(rows, Cin, device="cuda") * 0.5).to(dt)
w = (torch.randn(Cout, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).to(dt)
sc, bi = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
out = torch.zeros(rows if out_padded else B * H * W, Cout, device="cuda", dtype=dt)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    run_gemm(a, w, Cout, out, taps=9, cin=Cin, H=H, W=W, in_padded=True, out_padded=bool(out_padded), sample_rows=(H + 2) * (W + 2),
             scale=sc, bias=bi, act=L.ACT_RELU, impl=L.IMPL_TCGEN05, tile_cfg=cfg)


run(); run()
ts = []
for _ in range(5):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = sorted(ts)[len(ts) // 2]
print(f"B={B} {H}x{W} {Cin}->{Cout} cfg={cfg}: {us:.1f} us  {2 * B * H * W * Cout * 9 * Cin / us / 1e6:.1f} TF/s")
