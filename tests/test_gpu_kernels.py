"""GPU parity tests of the individual kernels, called through the C ABI.

Floating-point kernels are compared with a plain PyTorch fp32 restatement of the same op
(tolerance written at each assert); integer / index work (peaks, rasterised IoU, J flags) is
compared bit-for-bit with the CPU oracle.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from crog_b200 import _lib as L
from crog_b200 import synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import (compact_nhwc, conv_w, maxerr, pad_nhwc, relerr, run_gemm, uncompact, unpad)

DEV = "cuda"
BF = torch.bfloat16
IMPLS = [("simt_f32", torch.float32, L.IMPL_SIMT, 2e-5), ("simt_bf16", BF, L.IMPL_SIMT, 6e-3), ("tc_bf16", BF, L.IMPL_TCGEN05, 6e-3)]


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def test_library_and_device():
    lib = L.lib()
    assert lib.crog_abi_version() == L.ABI_VERSION
    assert lib.crog_check_device() == 0


@pytest.mark.parametrize("name,dt,impl,tol", IMPLS)
@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (200, 128, 512), (1088, 1536, 512), (300, 192, 2048), (130, 1024, 64)])
def test_plain_gemm(name, dt, impl, tol, M, N, K):
    a, w = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5)
    bias = _rand(N, seed=3)
    out = torch.zeros(M, N, device=DEV, dtype=dt)
    run_gemm(a.to(dt), w.to(dt), N, out, bias=bias, impl=impl)
    want = a.to(dt).float() @ w.to(dt).float().t() + bias
    assert relerr(out, want) < tol, name


@pytest.mark.parametrize("name,dt,impl,tol", IMPLS)
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 13, 13, 64, 64), (1, 26, 26, 128, 192), (3, 9, 20, 64, 128), (1, 5, 200, 64, 64),
                                            (2, 30, 30, 64, 32)])
def test_conv3x3_padded(name, dt, impl, tol, B, H, W, Cin, Cout):
    x, w = _rand(B, Cin, H, W, seed=4), _rand(Cout, Cin, 3, 3, seed=5, scale=(9 * Cin) ** -0.5)
    sc, bi = torch.rand(Cout, device=DEV) + 0.5, _rand(Cout, seed=6)
    want = F.relu(F.conv2d(x.to(dt).float(), w.to(dt).float(), padding=1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1))
    a = pad_nhwc(x, dt)
    # padded -> padded (halo must stay zero) and padded -> compact
    outp = torch.zeros(B * (H + 2) * (W + 2), Cout, device=DEV, dtype=dt)
    run_gemm(a, conv_w(w, dt), Cout, outp, taps=9, H=H, W=W, in_padded=True, out_padded=True, sample_rows=(H + 2) * (W + 2),
             scale=sc, bias=bi, act=L.ACT_RELU, impl=impl)
    assert relerr(unpad(outp, B, H, W), want) < tol, name
    full = outp.float().view(B, H + 2, W + 2, Cout)
    assert full[:, 0].abs().max() == 0 and full[:, :, 0].abs().max() == 0 and full[:, -1].abs().max() == 0
    outc = torch.zeros(B * H * W, Cout, device=DEV, dtype=dt)
    run_gemm(a, conv_w(w, dt), Cout, outc, taps=9, H=H, W=W, in_padded=True, out_padded=False, sample_rows=(H + 2) * (W + 2),
             scale=sc, bias=bi, act=L.ACT_RELU, impl=impl)
    assert relerr(uncompact(outc, B, H, W), want) < tol, name


@pytest.mark.parametrize("name,dt,impl,tol", IMPLS)
def test_gemm_epilogue_variants(name, dt, impl, tol):
    B, H, W, Cin, Cout = 2, 13, 13, 128, 256
    x, w = _rand(B, Cin, H, W, seed=7), _rand(Cout, Cin, 1, 1, seed=8, scale=Cin ** -0.5)
    a = compact_nhwc(x, dt)
    conv = F.conv2d(x.to(dt).float(), w.to(dt).float())
    sc, bi = torch.rand(Cout, device=DEV) + 0.5, _rand(Cout, seed=9)
    sc2, bi2 = torch.rand(Cout, device=DEV) + 0.5, _rand(Cout, seed=10)
    gate = torch.rand(B, Cout, device=DEV)
    # text gate (layers.py:376-379): relu(bn2(relu(bn1(conv)) * s)) into a padded output
    want = F.relu((F.relu(conv * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1)) * gate[:, :, None, None]) * sc2.view(1, -1, 1, 1) + bi2.view(1, -1, 1, 1))
    outp = torch.zeros(B * (H + 2) * (W + 2), Cout, device=DEV, dtype=dt)
    run_gemm(a, conv_w(w, dt), Cout, outp, H=H, W=W, out_padded=True, sample_rows=H * W, scale=sc, bias=bi, act=L.ACT_RELU,
             gate=gate, scale2=sc2, bias2=bi2, impl=impl)
    assert relerr(unpad(outp, B, H, W), want) < tol
    # periodic add-matrix + residual + relu, fp32 output stream, channel-slice output
    addm = _rand(H * W, Cout, seed=11)
    res = _rand(B * H * W, Cout + 64, seed=12)
    out = res.clone()
    run_gemm(a, conv_w(w, dt), Cout, out, H=H, W=W, sample_rows=H * W, addmat=addm, bias=bi, residual=out[:, 64:], residual_relu=True,
             impl=impl, out_col0=64)
    # residual pointer is a column slice of the same buffer (in place)
    want2 = F.relu(compact_nhwc(conv, torch.float32) + addm.repeat(B, 1) + bi + res[:, 64:])
    assert relerr(out[:, 64:], want2) < tol
    assert torch.equal(out[:, :64], res[:, :64])
    # QuickGELU
    out3 = torch.zeros(B * H * W, Cout, device=DEV, dtype=dt)
    run_gemm(a, conv_w(w, dt), Cout, out3, H=H, W=W, sample_rows=H * W, bias=bi, act=L.ACT_QUICKGELU, impl=impl)
    t = compact_nhwc(conv, torch.float32) + bi
    assert relerr(out3, t * torch.sigmoid(1.702 * t)) < tol


@pytest.mark.parametrize("name,dt,impl,tol", IMPLS)
def test_per_sample_weights_small_n(name, dt, impl, tol):
    """the projector's dynamic 3x3 conv: per-sample [16, 9*C] weights, fp32 compact output (layers.py:95-123)."""
    B, H, W, Cc = 3, 12, 10, 64
    x = _rand(B, Cc, H, W, seed=13)
    w = _rand(B, 16, Cc, 3, 3, seed=14, scale=(9 * Cc) ** -0.5)
    a = pad_nhwc(x, dt)
    wk = w.permute(0, 1, 3, 4, 2).reshape(B * 16, 9 * Cc).to(dt).contiguous()
    out = torch.zeros(B * H * W, 16, device=DEV, dtype=torch.float32)
    run_gemm(a, wk, 16, out, taps=9, H=H, W=W, in_padded=True, sample_rows=(H + 2) * (W + 2), w_sample_stride=16 * 9 * Cc, impl=impl)
    want = torch.stack([F.conv2d(x[b:b + 1].to(dt).float(), w[b].to(dt).float(), padding=1)[0] for b in range(B)])
    assert relerr(uncompact(out, B, H, W), want) < tol


@pytest.mark.parametrize("dt", [torch.float32, BF])
def test_resample_modes(dt):
    B, H, W, Cc = 2, 12, 10, 64
    x = _rand(B, Cc, H, W, seed=15)
    lib = L.lib()
    tol = 1e-6 if dt == torch.float32 else 5e-3
    xq = x.to(dt).float()
    for in_p in (False, True):
        src = pad_nhwc(x, dt) if in_p else compact_nhwc(x, dt)
        for mode, want in ((0, xq), (1, F.avg_pool2d(xq, 2)), (2, F.interpolate(xq, scale_factor=2, mode="bilinear"))):
            OH, OW = want.shape[-2:]
            for out_p in (False, True):
                ld = Cc + 64
                dst = torch.zeros(B * ((OH + 2) * (OW + 2) if out_p else OH * OW), ld, device=DEV, dtype=dt)
                L.check(lib.crog_resample(src.data_ptr(), Cc, int(in_p), dst.data_ptr() + 64 * dst.element_size(), ld, int(out_p),
                                          B, H, W, Cc, mode, L.dtype_code(dt), L.stream_ptr()))
                torch.cuda.synchronize()
                got = (unpad if out_p else uncompact)(dst, B, OH, OW)[:, 64:]
                assert relerr(got, want) < tol, (in_p, mode, out_p)
                assert dst[:, :64].abs().max() == 0


@pytest.mark.parametrize("dt", [torch.float32, BF])
def test_stem_conv1(dt):
    B, S = 2, 64
    img, w = _rand(B, 3, S, S, seed=16), _rand(32, 3, 3, 3, seed=17, scale=27 ** -0.5)
    sc, bi = torch.rand(32, device=DEV) + 0.5, _rand(32, seed=18)
    out = torch.full((B * (S // 2 + 2) ** 2, 64), 7.0, device=DEV, dtype=dt)
    out.view(B, S // 2 + 2, S // 2 + 2, 64)[:, 0] = 0; out.view(B, S // 2 + 2, S // 2 + 2, 64)[:, -1] = 0
    out.view(B, S // 2 + 2, S // 2 + 2, 64)[:, :, 0] = 0; out.view(B, S // 2 + 2, S // 2 + 2, 64)[:, :, -1] = 0
    out[:, 32:] = 0  # the padding channels belong to the caller (zero-initialised plan buffer); the kernel never writes them
    L.check(L.lib().crog_stem_conv1(img.data_ptr(), B, S, S, w.data_ptr(), sc.data_ptr(), bi.data_ptr(), 32, out.data_ptr(), 64,
                                    L.dtype_code(dt), 0, L.stream_ptr()))
    torch.cuda.synchronize()
    want = F.relu(F.conv2d(img, w, stride=2, padding=1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1))
    got = unpad(out, B, S // 2, S // 2)
    assert relerr(got[:, :32], want) < (1e-5 if dt == torch.float32 else 5e-3)
    assert got[:, 32:].abs().max() == 0


@pytest.mark.parametrize("D", [512, 2048])
def test_layernorm(D):
    rows = 77
    x, g, b, res = _rand(rows, D, seed=19, scale=3.0), torch.rand(D, device=DEV) + 0.5, _rand(D, seed=20), _rand(rows, D, seed=21)
    lib = L.lib()
    out = torch.zeros(rows, D, device=DEV)
    L.check(lib.crog_layernorm(x.data_ptr(), L.F32, g.data_ptr(), b.data_ptr(), None, out.data_ptr(), L.F32, rows, D, 1e-5, L.stream_ptr()))
    torch.cuda.synchronize()
    assert maxerr(out, F.layer_norm(x, (D,), g, b)) < 1e-4
    xb = x.to(BF)
    acc = res.clone()
    L.check(lib.crog_layernorm(xb.data_ptr(), L.BF16, g.data_ptr(), b.data_ptr(), acc.data_ptr(), acc.data_ptr(), L.F32, rows, D, 1e-5, L.stream_ptr()))
    torch.cuda.synchronize()
    assert maxerr(acc, res + F.layer_norm(xb.float(), (D,), g, b)) < 1e-4
    ob = torch.zeros(rows, D, device=DEV, dtype=BF)
    L.check(lib.crog_layernorm(x.data_ptr(), L.F32, g.data_ptr(), b.data_ptr(), None, ob.data_ptr(), L.BF16, rows, D, 1e-5, L.stream_ptr()))
    torch.cuda.synchronize()
    assert relerr(ob, F.layer_norm(x, (D,), g, b)) < 5e-3


def _ref_attention(q, k, v, heads, causal, pad):
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


@pytest.mark.parametrize("dt", [torch.float32, BF])
@pytest.mark.parametrize("Tq,Tk,heads,causal,padded", [(17, 17, 8, True, False), (169, 169, 4, False, False), (676, 676, 2, False, False),
                                                        (300, 20, 8, False, True)])
def test_attention(dt, Tq, Tk, heads, causal, padded):
    B, D = 2, heads * 64
    qkv = _rand(B * Tq, 3 * D, seed=22).to(dt)
    kv = _rand(B * Tk, 2 * D, seed=23).to(dt) if Tk != Tq else None
    word = None
    if padded:
        word = torch.zeros(B, Tk, dtype=torch.int64, device=DEV)
        word[0, :7] = 5; word[1, :13] = 9
    o = torch.zeros(B * Tq, D, device=DEV, dtype=dt)
    es = qkv.element_size()
    if kv is None:
        kp, vp, ldk = qkv.data_ptr() + D * es, qkv.data_ptr() + 2 * D * es, 3 * D
        kf, vf = qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        kp, vp, ldk = kv.data_ptr(), kv.data_ptr() + D * es, 2 * D
        kf, vf = kv[:, :D], kv[:, D:]
    L.check(L.lib().crog_attention(qkv.data_ptr(), 3 * D, kp, ldk, vp, ldk, o.data_ptr(), D, B, heads, Tq, Tk, 0.125, int(causal),
                                   word.data_ptr() if word is not None else None, L.dtype_code(dt), L.stream_ptr()))
    torch.cuda.synchronize()
    want = _ref_attention(qkv[:, :D].float().view(B, Tq, D), kf.float().reshape(B, Tk, D), vf.float().reshape(B, Tk, D), heads, causal, word)
    assert relerr(o.view(B, Tq, D), want) < (2e-5 if dt == torch.float32 else 6e-3)


def test_embed_gather_cast_split():
    lib = L.lib()
    B, Lt, D = 3, 17, 512
    _, word = synth.make_inputs(B, Lt, size=32)
    word = word.to(DEV)
    emb, pos = _rand(49408, D, seed=24), _rand(77, D, seed=25)
    x = torch.zeros(B * Lt, D, device=DEV)
    L.check(lib.crog_embed_tokens(word.data_ptr(), emb.data_ptr(), pos.data_ptr(), x.data_ptr(), B, Lt, D, emb.shape[0], L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(x.view(B, Lt, D), emb[word] + pos[:Lt])
    eot = torch.zeros(B, D, device=DEV, dtype=BF)
    L.check(lib.crog_gather_eot(word.data_ptr(), x.data_ptr(), L.F32, eot.data_ptr(), L.BF16, B, Lt, D, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(eot, x.view(B, Lt, D)[torch.arange(B), word.argmax(-1)].to(BF))
    y = torch.zeros(B * Lt * D, device=DEV, dtype=BF)
    L.check(lib.crog_cast(x.data_ptr(), L.F32, y.data_ptr(), L.BF16, x.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(y, x.flatten().to(BF))
    heads = _rand(1000, 16, seed=26)
    out = torch.zeros(5, 1000, device=DEV)
    L.check(lib.crog_split_heads(heads.data_ptr(), 16, out.data_ptr(), 1000, 5, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out, heads[:, :5].t().contiguous())


@pytest.mark.parametrize("dt", [torch.float32, BF])
def test_dynw_fold(dt):
    lib = L.lib()
    B, WD, Cc, NH, ZR, CP = 2, 1024, 256, 5, 48, 320
    state = _rand(B, WD, seed=27)
    tw, tb = _rand(9 * Cc + 1, WD, seed=28, scale=WD ** -0.5), _rand(9 * Cc + 1, seed=29)
    vw, vb = _rand(NH * Cc, Cc, seed=30, scale=Cc ** -0.5), _rand(NH * Cc, seed=31)
    scratch = torch.zeros(B, 9 * Cc + 1, device=DEV)
    wf = torch.full((B, ZR, CP), 3.0, device=DEV, dtype=dt)
    L.check(lib.crog_dynw_fold(state.data_ptr(), L.F32, tw.data_ptr(), tb.data_ptr(), vw.data_ptr(), vb.data_ptr(), scratch.data_ptr(),
                               wf.data_ptr(), L.dtype_code(dt), B, WD, Cc, NH, ZR, CP, L.stream_ptr()))
    torch.cuda.synchronize()
    t = state @ tw.t() + tb
    assert maxerr(scratch, t) < 1e-4
    wd = t[:, :-1].view(B, Cc, 9)  # [b, c, tap]
    V = vw.view(NH, Cc, Cc)  # [h, c, j]
    want = torch.zeros(B, NH, 9, CP, device=DEV)
    want[..., :Cc] = torch.einsum("bct,hcj->bhtj", wd, V)
    want[..., Cc] = torch.einsum("bct,hc->bht", wd, vb.view(NH, Cc))
    want[:, :, 4, Cc] += t[:, -1:].expand(B, NH)
    assert relerr(wf[:, :9 * NH], want.view(B, 9 * NH, CP)) < (1e-5 if dt == torch.float32 else 5e-3)
    assert (wf[:, 9 * NH:] == 3.0).all() and wf[:, :9 * NH, Cc + 1:].abs().max() == 0


def test_dynconv_gather_equals_grouped_conv():
    """Z = per-tap partial products in a zero-haloed matrix; the gather must equal a 3x3 pad-1 convolution."""
    lib = L.lib()
    B, H, W, NH, ZR = 2, 9, 7, 5, 48
    zt = _rand(B, NH, 9, H, W, seed=60)  # z[b, h, tap, y, x]
    z = torch.zeros(B, H + 2, W + 2, ZR, device=DEV)
    z[:, 1:-1, 1:-1, :9 * NH] = zt.permute(0, 3, 4, 1, 2).reshape(B, H, W, 9 * NH)
    out = torch.zeros(NH, B, H, W, device=DEV)
    L.check(lib.crog_dynconv_gather(z.data_ptr(), ZR, out.data_ptr(), B, H, W, NH, L.stream_ptr()))
    torch.cuda.synchronize()
    zp = F.pad(zt, (1, 1, 1, 1))
    want = sum(zp[:, :, t, t // 3:t // 3 + H, t % 3:t % 3 + W] for t in range(9)).permute(1, 0, 2, 3)
    assert maxerr(out, want) < 1e-5


def test_sigmoid_bicubic():
    from crog_b200.engine import postprocess

    maps = [_rand(3, 1, 104, 104, seed=32 + i, scale=3.0) for i in range(5)]
    got = postprocess(maps, (416, 416))
    torch.cuda.synchronize()
    for i, m in enumerate(maps):
        t = torch.sigmoid(m) if i in (0, 1, 4) else m
        want = F.interpolate(t, size=(416, 416), mode="bicubic", align_corners=True)[:, 0]
        assert maxerr(got[i], want) < 2e-5, i


# ------------------------------------------------------------------ tail: bit-exact against the oracle
def test_detect_grasps_unaligned_maps_and_odd_widths():
    """Maps that are planes of a larger tensor at an odd element offset, and widths that are no multiple of 4, take the
    generic scan kernel (the bulk-copy staged one needs 16-byte aligned rows): same peaks as the oracle."""
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail_c as TC

    for H, W, off in ((48, 64, 1), (37, 53, 0), (37, 53, 3)):
        q, s, c, w = synth.make_tail_maps(3, "blobs", seed=31 + off, size=64)
        q, s, c, w = [np.ascontiguousarray(a[:, :H, :W]) for a in (q, s, c, w)]
        dev = []
        for a in (q, s, c, w):
            buf = torch.zeros(a.size + 8, device=DEV)
            view = buf[off:off + a.size].view(a.shape)
            view.copy_(torch.from_numpy(a))
            dev.append(view)
        assert off == 0 or dev[0].data_ptr() % 16 != 0
        peaks, n, grasps = GE.detect_grasps_batched(*dev, 5)
        torch.cuda.synchronize()
        for b in range(3):
            g_ref, rc_ref = TC.detect_grasps(q[b], s[b], c[b], w[b], 5)
            assert int(n[b]) == len(rc_ref)
            assert np.array_equal(peaks[b, :int(n[b])].cpu().numpy(), rc_ref.astype(np.int32))


def _check_detect(q, s, c, w, K):
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail_c as TC

    peaks, n, grasps = GE.detect_grasps_batched(*[torch.from_numpy(a).to(DEV) for a in (q, s, c, w)], K)
    torch.cuda.synchronize()
    peaks, n, grasps = peaks.cpu().numpy(), n.cpu().numpy(), grasps.cpu().numpy()
    for b in range(q.shape[0]):
        g_ref, rc_ref = TC.detect_grasps(q[b], s[b], c[b], w[b], K)
        assert n[b] == len(rc_ref), (b, n[b], len(rc_ref))
        assert np.array_equal(peaks[b, :n[b]], rc_ref.astype(np.int32)), b
        assert (peaks[b, n[b]:] == -1).all()
        got = grasps[b, :n[b]]
        assert np.array_equal(got[:, :4], g_ref[:, :4]), b  # x, y, width*100, 20: bit-exact
        # angle: float32(atan2)/2 promoted to float64 -> allow 1 float32 ulp of the half-angle (A.3)
        ulp = np.spacing(np.abs(g_ref[:, 4] / 180 * np.pi).astype(np.float32)).astype(np.float64) * 180 / np.pi
        assert (np.abs(got[:, 4] - g_ref[:, 4]) <= ulp + 1e-12).all(), b
    return peaks, n, grasps


@pytest.mark.parametrize("kind,n,size", [("blobs", 6, 416), ("stress", 20, 416), ("blobs", 3, 100), ("stress", 3, 131)])
@pytest.mark.parametrize("K", [1, 5])
def test_detect_grasps_matches_oracle(kind, n, size, K):
    q, s, c, w = synth.make_tail_maps(n, kind, seed=40, size=size)
    _check_detect(q, s, c, w, K)


def test_detect_edge_cases():
    size = 64
    q = np.zeros((6, size, size), np.float32)
    q[0] = 0.7                                   # constant image: trivial -> no peaks
    q[1, 20:30, 20:40] = 0.9                     # big plateau: ties, min-distance suppression
    q[2, 1, 5] = 1.0; q[2, 30, 31] = np.float32(0.4); q[2, 40, 41] = np.nextafter(np.float32(0.4), np.float32(1))  # border, strict >
    q[3] = np.float32(0.5); q[3, 10, 10] = 0.49  # almost-constant plateau covering the map (exact fallback path)
    rng = np.random.default_rng(3)
    q[4] = np.floor(rng.random((size, size)) * 4).astype(np.float32) / 4  # heavy ties everywhere
    q[5, 2, 2] = 0.8; q[5, size - 3, size - 3] = 0.8; q[5, 2, size - 3] = 0.8  # first/last interior pixels
    s = rng.normal(size=q.shape).astype(np.float32); c = rng.normal(size=q.shape).astype(np.float32)
    w = rng.random(q.shape).astype(np.float32)
    for K in (1, 5, 9):
        _check_detect(q, s, c, w, K)


def test_detect_exact_fallback_on_wide_plateau():
    """More survivors requested than any warp segment keeps: the select kernel must notice and the exact
    sweep kernel must take over (non-square map, K=20)."""
    H, W = 16, 256
    q = np.full((2, H, W), 0.5, np.float32)
    q[0, 8, 100] = 0.25
    q[1] = np.linspace(0.41, 0.9, W, dtype=np.float32)[None, :].repeat(H, 0)  # ramps: one candidate column
    rng = np.random.default_rng(4)
    s = rng.normal(size=q.shape).astype(np.float32); c = rng.normal(size=q.shape).astype(np.float32)
    w = rng.random(q.shape).astype(np.float32)
    for K in (5, 20, 32):
        _check_detect(q, s, c, w, K)


def test_jaccard_matches_oracle():
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail_c as TC

    B, K, M = 24, 5, 64
    rng = np.random.default_rng(9)
    gt, cnt = synth.make_gt_rects(B, M, seed=4)
    grasps = np.zeros((B, K, 5), np.float64)
    n = rng.integers(0, K + 1, B).astype(np.int32)
    n[:4] = [0, 1, 5, 5]
    for b in range(B):
        for k in range(K):
            m = rng.integers(0, cnt[b])
            if k % 2 == 0:  # near a GT rectangle so overlaps are common
                grasps[b, k] = [gt[b, m, 0] + rng.uniform(-12, 12), gt[b, m, 1] + rng.uniform(-12, 12), rng.uniform(0, 105), 20,
                                gt[b, m, 4] + rng.uniform(-35, 35)]
            else:
                grasps[b, k] = [rng.uniform(0, 500), rng.uniform(0, 500), rng.uniform(0, 105), 20, rng.uniform(-90, 90)]
    gt_dev = torch.from_numpy(gt.copy()).to(DEV)
    counters = torch.zeros(4, dtype=torch.int64, device=DEV)
    flags, inter, uni = GE.jacquard_batched(torch.from_numpy(grasps).to(DEV), torch.from_numpy(n).to(DEV), gt_dev,
                                            torch.from_numpy(cnt).to(DEV), counters=counters, want_counts=True)
    torch.cuda.synchronize()
    flags, inter, uni, gt_after = flags.cpu().numpy(), inter.cpu().numpy(), uni.cpu().numpy(), gt_dev.cpu().numpy()
    c_ref = np.zeros(4, np.int64)
    for b in range(B):
        g_ref = gt[b, :cnt[b]].copy()
        j1 = TC.jacquard(grasps[b, :min(n[b], 1)], g_ref) if n[b] else 0
        jk = TC.jacquard(grasps[b, :n[b]], g_ref) if n[b] else 0
        if not n[b]:
            g_ref[:, 3] = 20; g_ref[:, 2] = np.clip(g_ref[:, 2], 0, 100)
        assert (flags[b, 0], flags[b, 1]) == (j1, jk), b
        assert np.array_equal(gt_after[b, :cnt[b]], g_ref), b  # in-place edit of the targets (grasp_eval.py:367-368)
        for k in range(n[b]):
            for m in range(cnt[b]):
                assert (inter[b, k, m], uni[b, k, m]) == TC.iou_counts(grasps[b, k], g_ref[m]), (b, k, m)
        c_ref += [j1, 1, jk, 1]
    assert np.array_equal(counters.cpu().numpy(), c_ref)


def test_reference_signature_wrappers():
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail as T

    q, s, c, w = synth.make_tail_maps(1, "blobs", seed=50, size=200)
    got, ang = GE.detect_grasps(q[0], s[0], c[0], w[0], 5)
    want, ang_ref = T.detect_grasps(q[0], s[0], c[0], w[0], 5)
    assert len(got) == len(want) and all(g[:4] == r[:4] for g, r in zip(got, want))
    assert np.abs(ang - ang_ref).max() <= 2e-7
    # calculate_iou uses the rectangle as given (no h:=20 edit), incl. a big one (generic path) and the x>=480 quirk
    for p, g in [([200, 200, 60, 20, 10], [205, 198, 80, 35, 20, 1]), ([240, 240, 400, 300, 17], [250, 230, 380, 20, 10, 1]),
                 ([470, 200, 90, 20, 0], [465, 205, 100, 20, 5, 1]), ([200, 200, 60, 20, 45], [200, 200, 60, 20, -5, 1])]:
        assert GE.calculate_iou(p, g) == T.calculate_iou(p, g), (p, g)
    gt = np.array([[200., 200., 150., 33., 5., 1.], [50., 60., 20., 20., 80., 1.]])
    assert GE.calculate_jacquard_index([[200., 200., 100., 20, 5.]], gt) == 1
    assert gt[0, 3] == 20 and gt[0, 2] == 100
    assert GE.calculate_jacquard_index([], np.array([[200., 200., 50., 20., 5., 1.]])) == 0
    assert GE.calculate_max_iou([[200, 200, 60, 20, 10]], [[200, 200, 60, 20, 10, 1]]) == 1.0


@pytest.mark.parametrize("case", ["conv3x3_pad2pad_tma", "conv3x3_pad2compact_legacy", "linear_residual_f32"])
def test_gemm_cta_pair_matches_single_cta(case, monkeypatch):
    """The cta_group::2 path (two CTAs share one 256 x 256 tile) must give the same bytes as the single-CTA tile loop:
    same k order, same fp32 accumulation, same epilogue code.  M is not a multiple of 256, so the last tile has a
    partly / fully out-of-range CTA."""
    torch.manual_seed(5)
    dt = torch.bfloat16
    if case.startswith("conv3x3"):
        B, H, W, Cin, Cout = 6, 58, 60, 256, 256  # M = 6*60*62 = 22320 padded rows (> 256*74, not a multiple of 256)
        x = torch.randn(B, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (Cin * 9) ** -0.5
        sc, bi = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
        a = pad_nhwc(x, dt)
        out_padded = case.endswith("tma")

        def run():
            rows = B * (H + 2) * (W + 2) if out_padded else B * H * W
            out = torch.zeros((rows, Cout), device="cuda", dtype=dt)
            run_gemm(a, conv_w(w, dt), Cout, out, taps=9, H=H, W=W, in_padded=True, out_padded=out_padded,
                     sample_rows=(H + 2) * (W + 2), scale=sc, bias=bi, act=L.ACT_RELU, impl=L.IMPL_TCGEN05)
            return out
    else:
        M, N, K = 256 * 80 + 77, 512, 2048
        a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
        bi = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")

        def run():
            out = res.clone()
            run_gemm(a, w, N, out, bias=bi, residual=out, impl=L.IMPL_TCGEN05)
            return out
    monkeypatch.setenv("CROG_GEMM_PAIR", "0")  # single-CTA tiles
    want = run()
    monkeypatch.delenv("CROG_GEMM_PAIR")       # default dispatch: CTA pairs for these shapes
    got = run()
    assert torch.equal(got, want), f"max diff {maxerr(got, want)}"


@pytest.mark.parametrize("case", ["conv3x3_pad2pad", "conv3x3_pad2compact", "conv3x3_cin64_n64", "linear_k512_addmat_bf16",
                                  "linear_k64_residual_relu", "linear_n128_f32_residual", "linear_ragged_n"])
def test_tile_cfgs_bit_identical(case):
    """Every CROG_TILE_* configuration the plan-time autotuner may pick (single CTA 128x64/128x128/128x256, CTA pairs
    256x256 / 256x128 with one or two epilogue groups, CONV3) accumulates the k-blocks in the same order through the same
    epilogue, so a forced configuration must reproduce the heuristic's bytes exactly; inapplicable ones must refuse."""
    torch.manual_seed(11)
    dt = torch.bfloat16
    if case.startswith("conv3x3"):
        if case == "conv3x3_cin64_n64":
            B, H, W, Cin, Cout = 3, 37, 41, 64, 64
        else:
            B, H, W, Cin, Cout = 4, 30, 33, 128, 256  # M = 4*32*35 = 4480 padded rows: ragged last pair tile
        x = torch.randn(B, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (Cin * 9) ** -0.5
        sc, bi = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
        a, wk = pad_nhwc(x, dt), conv_w(w, dt)
        out_padded = case != "conv3x3_pad2compact"

        def run(cfg, reverse=0):
            rows = B * (H + 2) * (W + 2) if out_padded else B * H * W
            out = torch.zeros((rows, Cout), device="cuda", dtype=dt)
            run_gemm(a, wk, Cout, out, taps=9, H=H, W=W, in_padded=True, out_padded=out_padded, reverse=reverse,
                     sample_rows=(H + 2) * (W + 2), scale=sc, bias=bi, act=L.ACT_RELU, impl=L.IMPL_TCGEN05, tile_cfg=cfg)
            return out
    else:
        M, N, K, odt = {"linear_k512_addmat_bf16": (676 * 5 + 3, 1536, 512, dt), "linear_k64_residual_relu": (5000, 256, 64, dt),
                        "linear_n128_f32_residual": (2704 + 129, 128, 512, torch.float32),
                        "linear_ragged_n": (1000, 328, 256, dt)}[case]
        a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
        bi = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda").to(odt)
        addmat = torch.randn(97, N, device="cuda") if "addmat" in case else None

        def run(cfg, reverse=0):
            out = res.clone()
            run_gemm(a, w, N, out, bias=bi, residual=None if addmat is not None else out, residual_relu="relu" in case,
                     addmat=addmat, sample_rows=97 if addmat is not None else 0, impl=L.IMPL_TCGEN05, tile_cfg=cfg, reverse=reverse)
            return out
    want = run(L.TILE_AUTO)
    ran = []
    for cfg in range(1, L.TILE_COUNT):
        try:
            got = run(cfg)
        except L.CrogError:
            continue  # does not apply to this shape
        ran.append(cfg)
        assert torch.equal(got, want), f"tile_cfg {cfg}: max diff {maxerr(got, want)}"
        # CrogGemm.reverse: the same tiles walked from the last to the first
        assert torch.equal(run(cfg, reverse=1), want), f"tile_cfg {cfg} reversed: max diff {maxerr(got, want)}"
    assert L.TILE_128x128 in ran and L.TILE_128x64 in ran
    if case != "conv3x3_cin64_n64":
        assert L.TILE_PAIR_256x128 in ran
    else:
        assert L.TILE_CONV3 in ran
    if case in ("conv3x3_pad2pad", "conv3x3_pad2compact"):  # activation-band configurations (3x3 on the padded layout)
        assert {L.TILE_BAND_PAIR_256x256, L.TILE_BAND_PAIR_256x256_E8, L.TILE_BAND_PAIR_256x128, L.TILE_BAND_128x128,
                L.TILE_BAND_128x256} <= set(ran)
    elif not case.startswith("conv3x3"):
        assert not any(c >= L.TILE_BAND_PAIR_256x256 for c in ran)


def test_plan_autotune_keeps_results():
    """The plan-time autotuner changes tile configurations only: maps before and after are bit-identical."""
    from crog_b200 import synth
    from crog_b200.model import CROG

    cfg = synth.default_cfg(17)
    model = CROG(cfg, precision="bf16", use_cuda_graph=False)
    model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed"))
    model = model.cuda()
    model.autotune = False
    img, word = synth.make_inputs(2, 17)
    want = torch.stack(model(img.cuda(), word.cuda())[0])
    plan = model.plan_for(2, 416)
    choice = plan.autotune(reps=2, min_gain=0.0)
    assert len(choice) == len(plan.gemm_ops)
    got = torch.stack(model(img.cuda(), word.cuda())[0])
    assert torch.equal(got, want)


@pytest.mark.parametrize("chunks", [1, 3, 4])
def test_decode_and_score_pipelined_equals_serial(chunks):
    """The two-stream, sub-batched tail issues the same kernels on slices: identical peaks, grasps, flags and counters."""
    from crog_b200 import synth
    from crog_b200.utils import grasp_eval as GE

    n = 10
    q, s, c, w = [torch.from_numpy(a).cuda() for a in synth.make_tail_maps(n, "blobs", 7, 160)]
    gt, cnt = synth.make_gt_rects(n, 64, seed=4)
    g1, g2 = torch.from_numpy(gt.copy()).cuda(), torch.from_numpy(gt.copy()).cuda()
    dcnt = torch.from_numpy(cnt).cuda()
    c1, c2 = torch.zeros(4, dtype=torch.int64, device="cuda"), torch.zeros(4, dtype=torch.int64, device="cuda")
    pk, npk, gr = GE.detect_grasps_batched(q, s, c, w, 5)
    fl = GE.jacquard_batched(gr, npk, g1, dcnt, counters=c1)
    pk2, npk2, gr2, fl2 = GE.decode_and_score_batched(q, s, c, w, g2, dcnt, 5, counters=c2, chunks=chunks)
    torch.cuda.synchronize()
    assert torch.equal(npk, npk2) and torch.equal(fl, fl2) and torch.equal(c1, c2) and torch.equal(g1, g2)
    for b in range(n):
        k = int(npk[b])
        assert torch.equal(pk[b, :k], pk2[b, :k]) and torch.equal(gr[b, :k], gr2[b, :k])


def test_attention_lazy_rescale_and_chained_layernorm():
    """(1) Logits that grow by far more than 2^8 from one key tile to the next force the in-TMEM rescale of the output
    row (csrc/attention_tc.cu) on every tile; logits that shrink leave the stale reference maximum in place.
    (2) crog_layernorm_chain equals two crog_layernorm calls."""
    B, heads, T = 2, 2, 500
    D = heads * 64
    for grow in (True, False):
        qkv = _rand(B * T, 3 * D, seed=31, scale=2.0)
        f = 1.0 + 4.0 * (torch.arange(T, device=DEV) // 128).float()
        if not grow:
            f = f.flip(0)
        qkv[:, D:2 * D] *= f.repeat(B)[:, None]
        qkv = qkv.to(BF)
        o = torch.zeros(B * T, D, device=DEV, dtype=BF)
        es = qkv.element_size()
        L.check(L.lib().crog_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + D * es, 3 * D, qkv.data_ptr() + 2 * D * es, 3 * D,
                                       o.data_ptr(), D, B, heads, T, T, 0.125, 0, None, L.BF16, L.stream_ptr()))
        torch.cuda.synchronize()
        want = _ref_attention(qkv[:, :D].float().view(B, T, D), qkv[:, D:2 * D].float().reshape(B, T, D),
                              qkv[:, 2 * D:].float().reshape(B, T, D), heads, False, None)
        assert torch.isfinite(o.float()).all()
        assert relerr(o.view(B, T, D), want) < 8e-3, grow
    rows, Dm = 1000, 512
    x = _rand(rows, Dm, seed=32).to(BF)
    res = _rand(rows, Dm, seed=33)
    g1, b1, g2, b2 = [_rand(Dm, seed=34 + i) for i in range(4)]
    y_ref, z_ref = res.clone(), torch.zeros(rows, Dm, device=DEV, dtype=BF)
    lib = L.lib()
    L.check(lib.crog_layernorm(x.data_ptr(), L.BF16, g1.data_ptr(), b1.data_ptr(), y_ref.data_ptr(), y_ref.data_ptr(), L.F32, rows, Dm, 1e-5, L.stream_ptr()))
    L.check(lib.crog_layernorm(y_ref.data_ptr(), L.F32, g2.data_ptr(), b2.data_ptr(), None, z_ref.data_ptr(), L.BF16, rows, Dm, 1e-5, L.stream_ptr()))
    y, z = res.clone(), torch.zeros(rows, Dm, device=DEV, dtype=BF)
    L.check(lib.crog_layernorm_chain(x.data_ptr(), L.BF16, g1.data_ptr(), b1.data_ptr(), y.data_ptr(), y.data_ptr(), g2.data_ptr(), b2.data_ptr(),
                                     z.data_ptr(), L.BF16, rows, Dm, 1e-5, L.stream_ptr()))
    torch.cuda.synchronize()
    assert maxerr(y, y_ref) <= 1e-6 * float(y_ref.abs().max()) and maxerr(z, z_ref) <= 2e-2


def test_layernorm_folded_into_gemm():
    """Linear -> ReLU -> LayerNorm -> Linear (decoder FFN, layers.py:302-308) with the LayerNorm folded into the second
    GEMM (row statistics emitted by the first GEMM's epilogue) against the unfused kernels and an fp32 torch reference."""
    torch.manual_seed(13)
    M, D, Fh = 676 * 3 + 5, 512, 2048
    dt = torch.bfloat16
    x = (torch.randn(M, D, device=DEV) * 0.7).to(dt)
    w0 = (torch.randn(Fh, D, device=DEV) * D ** -0.5)
    b0 = torch.randn(Fh, device=DEV) * 0.3
    gam, bet = torch.rand(Fh, device=DEV) + 0.5, torch.randn(Fh, device=DEV) * 0.2
    w4 = torch.randn(D, Fh, device=DEV) * Fh ** -0.5
    b4 = torch.randn(D, device=DEV) * 0.1
    res = torch.randn(M, D, device=DEV)
    # fp32 reference on the bf16-rounded operands
    h = torch.relu(x.float() @ w0.to(dt).float().t() + b0)
    want = torch.nn.functional.layer_norm(h, (Fh,), gam, bet, 1e-5) @ w4.t() + b4 + res

    def gemm(a, w, N, out, **kw):
        g = L.CrogGemm()
        g.a, g.a_rows, g.a_ld, g.cin, g.taps, g.dtype, g.M = a.data_ptr(), a.shape[0], a.shape[1], a.shape[1], 1, L.BF16, a.shape[0]
        g.w, g.N, g.out, g.out_ld, g.out_dtype, g.impl = w.data_ptr(), N, out.data_ptr(), out.shape[1], L.dtype_code(out.dtype), L.IMPL_TCGEN05
        for k, v in kw.items():
            setattr(g, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))

    # unfused: GEMM -> LayerNorm kernel -> GEMM
    ff = torch.zeros(M, Fh, device=DEV, dtype=dt); ff2 = torch.zeros_like(ff)
    gemm(x, w0.to(dt), Fh, ff, bias=b0, act=L.ACT_RELU)
    L.check(L.lib().crog_layernorm(ff.data_ptr(), L.BF16, gam.data_ptr(), bet.data_ptr(), None, ff2.data_ptr(), L.BF16, M, Fh, 1e-5, L.stream_ptr()))
    out_a = res.clone()
    gemm(ff2, w4.to(dt), D, out_a, bias=b4, residual=out_a, res_ld=D)
    # folded
    stats = torch.zeros(M, Fh // 64, 2, device=DEV)
    ffb = torch.zeros(M, Fh, device=DEV, dtype=dt)
    gemm(x, w0.to(dt), Fh, ffb, bias=b0, act=L.ACT_RELU, row_stats_out=stats, row_stats_chunks=Fh // 64, row_stats_width=Fh, row_stats_eps=1e-5)
    w4g = (w4 * gam[None, :]).to(dt)
    s_vec, c_vec = w4g.float().sum(1).contiguous(), (w4 @ bet + b4).contiguous()
    out_b = res.clone()
    gemm(ffb, w4g, D, out_b, scale=s_vec, bias=c_vec, residual=out_b, res_ld=D, row_stats_in=stats, row_stats_chunks=Fh // 64,
         row_stats_width=Fh, row_stats_eps=1e-5)
    torch.cuda.synchronize()
    assert torch.equal(ff, ffb)  # emitting the statistics does not change the producer's output
    hs = ff.float()
    assert torch.allclose(stats[..., 0].sum(1), hs.sum(1), rtol=2e-3, atol=2e-2) and torch.allclose(stats[..., 1].sum(1), (hs * hs).sum(1), rtol=4e-3)
    ea, eb = relerr(out_a - res, want - res), relerr(out_b - res, want - res)
    assert ea < 1e-2 and eb < 1e-2, (ea, eb)
    assert eb < 1.5 * ea + 1e-3, (ea, eb)  # the fold is as accurate as the separate LayerNorm pass


def test_gemm_second_operand_equals_two_gemms():
    """CrogGemm.a2: out = act([a | a2] [w1 | w2]^T + b) in one launch against the two contractions done separately in fp32
    (the fused conv3 + downsample of a bottleneck, clip.py:44-57), for a single-CTA and a CTA-pair configuration."""
    torch.manual_seed(17)
    M, K1, K2, N = 2704 * 2 + 37, 128, 512, 512
    dt = torch.bfloat16
    a1 = (torch.randn(M, K1, device=DEV) * 0.5).to(dt)
    a2 = (torch.randn(M, K2, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K1 + K2, device=DEV) * (K1 + K2) ** -0.5).to(dt)
    b = torch.randn(N, device=DEV)
    want = torch.relu(a1.float() @ w[:, :K1].float().t() + a2.float() @ w[:, K1:].float().t() + b)
    outs = []
    for cfg in (L.TILE_AUTO, L.TILE_128x128, L.TILE_PAIR_256x256_E8, L.TILE_128x64):
        out = torch.zeros(M, N, device=DEV, dtype=dt)
        g = L.CrogGemm()
        g.a, g.a_rows, g.a_ld, g.cin, g.taps, g.dtype, g.M = a1.data_ptr(), M, K1, K1, 1, L.BF16, M
        g.a2, g.a2_ld, g.cin2 = a2.data_ptr(), K2, K2
        g.w, g.N, g.bias, g.act = w.data_ptr(), N, b.data_ptr(), L.ACT_RELU
        g.out, g.out_ld, g.out_dtype, g.impl, g.tile_cfg = out.data_ptr(), N, L.BF16, L.IMPL_TCGEN05, cfg
        L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))
        torch.cuda.synchronize()
        assert relerr(out, want) < 4e-3, cfg
        outs.append(out)
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    g.impl = L.IMPL_SIMT  # the CUDA-core path does not implement it and must say so
    with pytest.raises(L.CrogError):
        L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))


# ------------------------------------------------------------------ round 2: SM budgets, masked taps, pixel-pair stem
@pytest.mark.parametrize("M,N,K,cap", [(1088, 1536, 512, 16), (43264, 256, 1024, 132), (43264, 512, 2048, 30), (300, 64, 64, 1)])
def test_gemm_grid_cap_is_bit_identical(M, N, K, cap):
    """CrogGemm.max_ctas only bounds the persistent grid: every tile is computed the same way by whichever CTA gets it."""
    a, w = _rand(M, K, seed=1).to(BF), _rand(N, K, seed=2, scale=K ** -0.5).to(BF)
    bias = _rand(N, seed=3)
    o0, o1 = torch.zeros(M, N, device=DEV, dtype=BF), torch.zeros(M, N, device=DEV, dtype=BF)
    run_gemm(a, w, N, o0, bias=bias, act=L.ACT_RELU, impl=L.IMPL_TCGEN05)
    run_gemm(a, w, N, o1, bias=bias, act=L.ACT_RELU, impl=L.IMPL_TCGEN05, max_ctas=cap)
    assert torch.equal(o0, o1)
    assert relerr(o0, F.relu(a.float() @ w.float().t() + bias)) < 6e-3


def test_conv3_tap_mask_skips_zero_taps():
    """Resident-weight 3x3 path with tap_mask: the masked taps (zero weights) are not contracted; result == all nine taps."""
    B, H, W, Cin, Cout = 2, 20, 24, 64, 64
    x = _rand(B, Cin, H, W, seed=4)
    wgt = _rand(Cout, Cin, 3, 3, seed=5, scale=(9 * Cin) ** -0.5)
    for keep_kx in ((0, 1), (1, 2)):
        wz = wgt.clone()
        mask = 0
        for kx in range(3):
            if kx not in keep_kx:
                wz[:, :, :, kx] = 0
            else:
                for ky in range(3):
                    mask |= 1 << (ky * 3 + kx)
        a = pad_nhwc(x, BF)
        o_all = torch.zeros(B * (H + 2) * (W + 2), Cout, device=DEV, dtype=BF)
        o_msk = torch.zeros_like(o_all)
        run_gemm(a, conv_w(wz, BF), Cout, o_all, taps=9, cin=Cin, H=H, W=W, in_padded=True, out_padded=True,
                 sample_rows=(H + 2) * (W + 2), impl=L.IMPL_TCGEN05)
        run_gemm(a, conv_w(wz, BF), Cout, o_msk, taps=9, cin=Cin, H=H, W=W, in_padded=True, out_padded=True,
                 sample_rows=(H + 2) * (W + 2), impl=L.IMPL_TCGEN05, tap_mask=mask)
        want = F.conv2d(x.to(BF).float(), wz.to(BF).float(), padding=1)
        assert relerr(unpad(o_msk, B, H, W), want) < 6e-3
        assert relerr(unpad(o_msk, B, H, W), unpad(o_all, B, H, W)) < 1e-3  # same products, fewer zero terms


def test_stem_conv1_pixel_pair_layout():
    B, S = 2, 64
    img = _rand(B, 3, S, S, seed=6)
    w, sc, bi = _rand(32, 3, 3, 3, seed=7, scale=0.3), torch.rand(32, device=DEV) + 0.5, _rand(32, seed=8, scale=0.1)
    OH, OWp = S // 2, S // 4
    out = torch.zeros(B * (OH + 2) * (OWp + 2), 64, device=DEV, dtype=BF)
    L.check(L.lib().crog_stem_conv1(img.data_ptr(), B, S, S, w.data_ptr(), sc.data_ptr(), bi.data_ptr(), 32, out.data_ptr(), 64,
                                    L.dtype_code(BF), 1, L.stream_ptr()))
    torch.cuda.synchronize()
    want = F.relu(F.conv2d(img, w, stride=2, padding=1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1))  # B,32,OH,OW
    grid = out.float().view(B, OH + 2, OWp + 2, 2, 32)
    got = grid[:, 1:-1, 1:-1].reshape(B, OH, OWp * 2, 32).permute(0, 3, 1, 2)
    assert maxerr(got, want) < 2e-2
    halo = grid.clone(); halo[:, 1:-1, 1:-1] = 0
    assert float(halo.abs().max()) == 0.0
