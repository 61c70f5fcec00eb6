"""Profiling helper (not a test): one 1x1-convolution / linear GEMM of a given shape with the bottleneck epilogue
(folded BN + residual + ReLU, bf16 out), under the heuristic tile choice or a forced CROG_TILE_*.
usage: [ncu ...] python tests/prof_gemm_shape.py M N K [tile_cfg] [residual=1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import run_gemm  # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
use_res = int(sys.argv[5]) if len(sys.argv) > 5 else 1
torch.manual_seed(0)
dt = torch.bfloat16
a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
sc, bi = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda").to(dt)
out = torch.zeros(M, N, device="cuda", dtype=dt)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    run_gemm(a, w, N, out, scale=sc, bias=bi, residual=res if use_res else None, residual_relu=bool(use_res), impl=L.IMPL_TCGEN05, tile_cfg=cfg)


run(); run()
ts = []
for _ in range(5):
    flush.zero_()  # operands come from HBM, as in the forward
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
by = (M * K + M * N * (2 if use_res else 1) + N * K) * 2
print(f"M={M} N={N} K={K} cfg={cfg} res={use_res}: {us:.1f} us  {2 * M * N * K / us / 1e6:.1f} TF/s  {by / us / 1e3:.1f} GB/s algorithmic")
