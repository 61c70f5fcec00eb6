"""Randomised sweep of crog_gemm (tcgen05 path, linear form): random M / N / K, bias, ReLU, residual, output dtype, walk
direction; every applicable tile configuration must reproduce the heuristic's bytes and agree with an fp32 matmul of the
bf16 operands.  python scripts/fuzz_gemm.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from crog_b200 import _lib as L
from gpu_util import run_gemm, relerr

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad, t0, nrun = 0, time.time(), 0
for t in range(trials):
    M = int(rng.choice([1, 7, 127, 128, 129, 255, 1088, 2705, 10816, 43264, int(rng.integers(1, 60000))]))
    N = int(rng.choice([8, 24, 64, 72, 128, 200, 256, 328, 512, 1536, 2048]))
    K = int(rng.choice([64, 128, 192, 256, 512, 1024, 2048]))
    odt = torch.float32 if rng.random() < 0.25 else torch.bfloat16
    relu, use_res, use_bias, use_scale = rng.random() < 0.5, rng.random() < 0.4, rng.random() < 0.8, rng.random() < 0.3
    torch.manual_seed(int(rng.integers(1 << 30)))
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") if use_bias else None
    scale = (torch.rand(N, device="cuda") + 0.5) if use_scale else None
    res = torch.randn(M, N, device="cuda").to(odt) if use_res else None
    want = a.float() @ w.float().t()
    if scale is not None: want = want * scale
    if bias is not None: want = want + bias
    if relu and not use_res: want = torch.relu(want)
    if use_res: want = want + res.float()
    if relu and use_res: want = torch.relu(want)

    def run(cfg, reverse):
        out = res.clone() if use_res else torch.zeros(M, N, device="cuda", dtype=odt)
        run_gemm(a, w, N, out, scale=scale, bias=bias, act=L.ACT_RELU if (relu and not use_res) else 0,
                 residual=out if use_res else None, residual_relu=bool(relu and use_res), impl=L.IMPL_TCGEN05, tile_cfg=cfg, reverse=reverse)
        return out
    try:
        base = run(0, 0)
    except L.CrogError as e:
        print("heuristic refused", dict(M=M, N=N, K=K), str(e)[:80]); bad += 1; continue
    e = relerr(base, want)
    if not e < 8e-3:
        bad += 1; print("NUMERIC", dict(M=M, N=N, K=K, odt=str(odt), relu=relu, res=use_res), e); continue
    for cfg in range(1, L.TILE_COUNT):
        for rev in (0, 1):
            try:
                got = run(cfg, rev)
            except L.CrogError:
                break
            nrun += 1
            if not torch.equal(got, base):
                bad += 1; print("TILE MISMATCH", dict(M=M, N=N, K=K, cfg=cfg, rev=rev, odt=str(odt), relu=relu, res=use_res)); break
print(f"{trials} trials, {nrun} forced-configuration runs compared, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
