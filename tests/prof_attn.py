"""Profiling helper (not a test): decoder self-attention shape (B=64, 8 heads, 676 tokens) and the attention pool
(B=64, 32 heads, 169 tokens) on the tcgen05 attention kernel.
usage: ncu --set full --import-source on -k regex:attention_tc ... python tests/prof_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L  # noqa: E402

torch.manual_seed(0)
lib = L.lib()
for B, heads, T in ((64, 8, 676), (64, 32, 169)):
    D = heads * 64
    qkv = (torch.randn(B * T, 3 * D, device="cuda")).to(torch.bfloat16)
    o = torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)

    def run():
        L.check(lib.crog_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D, o.data_ptr(), D,
                                   B, heads, T, T, 0.125, 0, None, L.BF16, L.stream_ptr()))

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
    gf = 4 * B * heads * T * T * 64 / 1e9
    print(f"B={B} heads={heads} T={T}: {ms * 1e3:.1f} us  {gf / ms:.1f} TF/s")
