"""Randomised sweep of crog_attention (bf16: the tcgen05 kernel for long sequences, the small-sequence kernel for short ones)
against torch softmax attention in fp32: random batch, heads, query / key lengths (ragged last tiles in both directions,
self and cross form), causal masks on self-attention, key-padding masks, and logit scales that force the in-TMEM rescale.
python scripts/fuzz_attention.py [trials] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import _lib as L

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def ref(q, k, v, heads, causal, pad):
    B, Tq, D = q.shape
    Tk = k.shape[1]
    Q = q.view(B, Tq, heads, 64).transpose(1, 2) * 0.125
    K = k.view(B, Tk, heads, 64).transpose(1, 2)
    V = v.view(B, Tk, heads, 64).transpose(1, 2)
    s = Q @ K.transpose(-1, -2)
    if causal:
        s = s + torch.full((Tq, Tk), float("-inf"), device=q.device).triu_(1)
    if pad is not None:
        s = s.masked_fill((pad == 0)[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(B, Tq, D)


bad, t0 = 0, time.time()
for t in range(trials):
    B = int(rng.integers(1, 5)); heads = int(rng.choice([1, 2, 4, 8]))
    Tq = int(rng.choice([1, 5, 17, 20, 77, 127, 128, 129, 169, 300, 500, 676, int(rng.integers(1, 900))]))
    cross = rng.random() < 0.4
    Tk = int(rng.choice([1, 3, 17, 20, 33, 128, 200])) if cross else Tq
    causal = (not cross) and rng.random() < 0.3
    padded = cross and rng.random() < 0.6
    scale = float(rng.choice([1.0, 1.0, 3.0, 6.0]))
    D = heads * 64
    torch.manual_seed(int(rng.integers(1 << 30)))
    qkv = (torch.randn(B * Tq, 3 * D, device="cuda") * scale).to(torch.bfloat16)
    kv = (torch.randn(B * Tk, 2 * D, device="cuda") * scale).to(torch.bfloat16) if cross else None
    word = None
    if padded:
        word = torch.zeros(B, Tk, dtype=torch.int64, device="cuda")
        for b in range(B):
            word[b, :int(rng.integers(1, Tk + 1))] = 7
    o = torch.zeros(B * Tq, D, device="cuda", dtype=torch.bfloat16)
    es = 2
    if kv is None:
        kp, vp, ldk = qkv.data_ptr() + D * es, qkv.data_ptr() + 2 * D * es, 3 * D
        kf, vf = qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        kp, vp, ldk = kv.data_ptr(), kv.data_ptr() + D * es, 2 * D
        kf, vf = kv[:, :D], kv[:, D:]
    L.check(L.lib().crog_attention(qkv.data_ptr(), 3 * D, kp, ldk, vp, ldk, o.data_ptr(), D, B, heads, Tq, Tk, 0.125, int(causal),
                                   word.data_ptr() if word is not None else None, L.BF16, L.stream_ptr()))
    torch.cuda.synchronize()
    want = ref(qkv[:, :D].float().view(B, Tq, D), kf.float().reshape(B, Tk, D), vf.float().reshape(B, Tk, D), heads, causal, word)
    e = float((o.view(B, Tq, D).float() - want).norm() / (want.norm() + 1e-12))
    if not (e < 1e-2 and bool(torch.isfinite(o.float()).all())):
        bad += 1
        print("MISMATCH", dict(B=B, heads=heads, Tq=Tq, Tk=Tk, causal=causal, padded=padded, scale=scale), e)
print(f"{trials} trials, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
