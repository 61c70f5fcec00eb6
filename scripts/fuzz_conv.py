"""Randomised sweep of crog_gemm as a 3x3 convolution on the zero-haloed NHWC layout: random batch / extent / channels,
padded or compact output, BN scale + bias + ReLU; every applicable tile configuration (per-tap, CTA pair, activation band,
resident-weight CONV3 and its variants) in both walk directions must reproduce the heuristic's bytes and agree with
F.conv2d on the bf16 operands.  python scripts/fuzz_conv.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from crog_b200 import _lib as L
from gpu_util import run_gemm, relerr, pad_nhwc, conv_w, unpad, uncompact

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad, t0, nrun = 0, time.time(), 0
dt = torch.bfloat16
for t in range(trials):
    B = int(rng.integers(1, 5)); H = int(rng.integers(3, 60)); W = int(rng.integers(3, 60))
    Cin = int(rng.choice([64, 64, 128, 256, 512])); Cout = int(rng.choice([32, 64, 64, 128, 256, 512]))
    out_padded = bool(rng.random() < 0.5)
    torch.manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (Cin * 9) ** -0.5
    sc, bi = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
    a, wk = pad_nhwc(x, dt), conv_w(w, dt)
    want = torch.relu(F.conv2d(x.to(dt).float(), w.to(dt).float(), padding=1) * sc[None, :, None, None] + bi[None, :, None, None])

    def run(cfg, rev):
        rows = B * (H + 2) * (W + 2) if out_padded else B * H * W
        out = torch.zeros((rows, Cout), device="cuda", dtype=dt)
        run_gemm(a, wk, Cout, out, taps=9, H=H, W=W, in_padded=True, out_padded=out_padded, reverse=rev,
                 sample_rows=(H + 2) * (W + 2), scale=sc, bias=bi, act=L.ACT_RELU, impl=L.IMPL_TCGEN05, tile_cfg=cfg)
        return out
    base = run(0, 0)
    got = unpad(base, B, H, W) if out_padded else uncompact(base, B, H, W)
    e = relerr(got, want)
    if not e < 8e-3:
        bad += 1; print("NUMERIC", dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, padded=out_padded), e); continue
    if out_padded:  # the halo of a padded output must be zero
        full = base.float().view(B, H + 2, W + 2, Cout)
        if float(full[:, 0].abs().max()) or float(full[:, -1].abs().max()) or float(full[:, :, 0].abs().max()) or float(full[:, :, -1].abs().max()):
            bad += 1; print("HALO", dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout)); continue
    for cfg in range(1, L.TILE_COUNT):
        for rev in (0, 1):
            try:
                o = run(cfg, rev)
            except L.CrogError:
                break
            nrun += 1
            if not torch.equal(o, base):
                bad += 1; print("TILE MISMATCH", dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, padded=out_padded, cfg=cfg, rev=rev)); break
print(f"{trials} trials, {nrun} forced-configuration runs compared, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
