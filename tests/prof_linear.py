"""Profiling helper (not a test): one linear GEMM out = relu?(a @ w^T + bias) at a decoder shape.
usage: python tests/prof_linear.py M N K [tile_cfg] [residual]   (ncu --profile-from-start off brackets the timed launches)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import run_gemm  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
res = len(sys.argv) > 5 and sys.argv[5] == "residual"
add = len(sys.argv) > 5 and sys.argv[5] == "addmat"  # periodic fp32 bias matrix (676 rows), as the decoder's in_proj has
torch.manual_seed(0)
dt = torch.bfloat16
a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
bias = torch.randn(N, device="cuda")
out = torch.zeros((M, N), device="cuda", dtype=dt)
addmat = torch.randn(676, N, device="cuda") if add else None


def run():
    run_gemm(a, w, N, out, bias=bias, residual=out if res else None, addmat=addmat, sample_rows=676 if add else 0,
             impl=L.IMPL_TCGEN05, tile_cfg=cfg)


run(); run()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
torch.cuda.profiler.start()
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g = L.CrogGemm()
    e0.record(); 
    # launch without the helper's synchronize so that the events bracket the kernel only
    import ctypes as C
    g.a, g.a_rows, g.a_ld, g.cin, g.taps, g.dtype, g.M = a.data_ptr(), M, K, K, 1, L.BF16, M
    g.w, g.N, g.bias = w.data_ptr(), N, bias.data_ptr()
    if res:
        g.residual, g.res_ld = out.data_ptr(), N
    if add:
        g.addmat, g.addmat_rows, g.sample_rows = addmat.data_ptr(), 676, 676
    g.out, g.out_ld, g.out_dtype, g.impl, g.tile_cfg = out.data_ptr(), N, L.BF16, L.IMPL_TCGEN05, cfg
    L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
torch.cuda.profiler.stop()
ts.sort()
print(f"M {M} N {N} K {K} cfg {cfg} residual {res} addmat {add}: {ts[len(ts)//2]:.1f} us  {2.0 * M * N * K / ts[len(ts)//2] / 1e6:.0f} TF/s")
