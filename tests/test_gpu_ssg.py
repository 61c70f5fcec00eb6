"""GPU parity tests of the SSG path (BASELINE config 4), through the C ABI.

Floating-point stages are compared with the CPU oracle (oracle/ssg_forward.py, pinned against the real reference by
oracle/make_golden_ssg.py) at the tolerance written at each assert; discrete decisions (kept anchors, Fast-NMS
survivors, class ids, peaks / decoded grasps given identical maps) and the Gaussian (float64 accumulation in scipy's
order) are compared bit-for-bit.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from crog_b200 import _lib as L
from crog_b200 import synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import compact_nhwc, pad_nhwc, relerr, run_gemm, uncompact, unpad

DEV = "cuda"
BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(DEV)


@pytest.mark.parametrize("dt", [torch.float32, BF])
@pytest.mark.parametrize("S,cin,Kp", [(36, 4, 256), (150, 4, 256), (150, 3, 192), (274, 4, 200)])
def test_stem7_patches(dt, S, cin, Kp):
    """7x7 / stride 2 / pad 3 patch rows, (tap, channel) order, zero-padded to Kp; several 64-pixel row segments, a ragged
    last segment, and the 3-channel (no depth) form."""
    B, O = 2, (S - 1) // 2 + 1
    rgb, depth = torch.rand(B, 3, S, S, device=DEV), torch.rand(B, 1, S, S, device=DEV)
    out = torch.full((B * O * O, Kp), 7.0, device=DEV, dtype=dt)
    L.check(L.lib().crog_stem7_patches(rgb.data_ptr(), depth.data_ptr() if cin == 4 else None, B, S, S, cin, Kp, out.data_ptr(),
                                       L.dtype_code(dt), L.stream_ptr()))
    torch.cuda.synchronize()
    img = (torch.cat([rgb, depth], 1) if cin == 4 else rgb).to(dt).float()
    cols = F.unfold(img, 7, padding=3, stride=2)  # [B, cin*49, L], channel-major rows
    want = cols.view(B, cin, 49, -1).permute(0, 3, 2, 1).reshape(B * O * O, 49 * cin)  # -> (tap, channel)
    assert torch.equal(out[:, :49 * cin].float(), want)
    assert out[:, 49 * cin:].abs().max() == 0


@pytest.mark.parametrize("dt", [torch.float32, BF])
def test_maxpool_patches_resample(dt):
    lib = L.lib()
    B, H, W, C = 2, 13, 10, 64
    x = _rand(B, C, H, W, seed=1)
    xq = x.to(dt).float()
    code = L.dtype_code(dt)
    for in_p in (False, True):
        src = pad_nhwc(x, dt) if in_p else compact_nhwc(x, dt)
        # MaxPool2d(3, 2, 1)
        want = F.max_pool2d(xq, 3, 2, 1)
        OH, OW = want.shape[-2:]
        for out_p in (False, True):
            dst = torch.zeros(B * ((OH + 2) * (OW + 2) if out_p else OH * OW), C, device=DEV, dtype=dt)
            L.check(lib.crog_maxpool3s2(src.data_ptr(), C, int(in_p), dst.data_ptr(), C, int(out_p), B, H, W, C, code, L.stream_ptr()))
            torch.cuda.synchronize()
            assert torch.equal((unpad if out_p else uncompact)(dst, B, OH, OW), want)
        # x[::2, ::2] and bilinear x2 with align_corners=True
        for mode, want in ((3, xq[:, :, ::2, ::2]), (4, F.interpolate(xq, scale_factor=2, mode="bilinear", align_corners=True))):
            OH, OW = want.shape[-2:]
            dst = torch.zeros(B * (OH + 2) * (OW + 2), C, device=DEV, dtype=dt)
            L.check(lib.crog_resample(src.data_ptr(), C, int(in_p), dst.data_ptr(), C, 1, B, H, W, C, mode, code, L.stream_ptr()))
            torch.cuda.synchronize()
            assert relerr(unpad(dst, B, OH, OW), want) < (1e-6 if dt == torch.float32 else 5e-3), mode
    # 3x3 / stride 2 / pad 1 patches of the padded layout
    src = pad_nhwc(x, dt)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.zeros(B * OH * OW, 9 * C, device=DEV, dtype=dt)
    L.check(lib.crog_patches3(src.data_ptr(), C, out.data_ptr(), B, H, W, C, 2, code, L.stream_ptr()))
    torch.cuda.synchronize()
    cols = F.unfold(xq, 3, padding=1, stride=2).view(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B * OH * OW, 9 * C)
    assert torch.equal(out.float(), cols)


def test_gemm_tanh_and_level_concat():
    """tanh epilogue and out_sample_rows: two 'levels' of one sample land back to back in one [B, rows, N] tensor."""
    B, Cc, N = 2, 64, 96
    levels = [(6, 5), (3, 3)]
    tot = sum(h * w for h, w in levels)
    out = torch.zeros(B * tot, N, device=DEV, dtype=torch.float32)
    wgt = _rand(N, Cc, 3, 3, seed=3) * (9 * Cc) ** -0.5
    bias = _rand(N, seed=4)
    off, wants = 0, []
    for i, (h, w) in enumerate(levels):
        x = _rand(B, Cc, h, w, seed=5 + i)
        a = pad_nhwc(x, BF)
        g = L.CrogGemm()
        wk = wgt.permute(0, 2, 3, 1).reshape(N, -1).to(BF).contiguous()
        g.a, g.a_rows, g.a_ld, g.cin, g.taps, g.dtype = a.data_ptr(), a.shape[0], Cc, Cc, 9, L.BF16
        g.M, g.sample_rows, g.H, g.W, g.in_padded, g.out_padded = a.shape[0], (h + 2) * (w + 2), h, w, 1, 0
        g.w, g.N, g.bias, g.act = wk.data_ptr(), N, bias.data_ptr(), L.ACT_TANH
        g.out, g.out_ld, g.out_dtype, g.out_sample_rows = out.data_ptr() + off * N * 4, N, L.F32, tot
        import ctypes as C
        L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))
        torch.cuda.synchronize()
        wants.append(torch.tanh(F.conv2d(x.to(BF).float(), wgt.to(BF).float(), bias, padding=1)).permute(0, 2, 3, 1).reshape(B, h * w, N))
        off += h * w
    want = torch.cat(wants, 1)
    assert relerr(out.view(B, tot, N), want) < 6e-3


def _ssg_pair(size, batch, mode, precision):
    from crog_b200.model import SSG
    from oracle import ssg_forward as O

    cfg = synth.ssg_cfg(img_size=size)
    sd = synth.make_ssg_state_dict(cfg, 0, mode)
    rgb, depth = synth.make_ssg_inputs(batch, size)
    model = SSG(cfg, precision=precision)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    got = model({"rgb": rgb.cuda(), "depth": depth.cuda()})
    torch.cuda.synchronize()
    ref = O.ssg_forward(sd, cfg, rgb, depth)
    return got, ref, model


KEYS = ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")


def test_ssg_forward_fp32_matches_oracle():
    got, ref, _ = _ssg_pair(288, 2, "perturbed", "fp32")
    assert got["anchors"] == ref["anchors"]
    for k in KEYS:
        assert got[k].shape == ref[k].shape, k
        err = float((got[k].cpu() - ref[k]).abs().max())
        assert err <= 1e-3, (k, err)  # fp32 mode: 1e-3 max-abs (north star); measured ~1e-5


def test_ssg_forward_bf16_tolerance_and_golden():
    got, ref, _ = _ssg_pair(288, 2, "perturbed", "bf16")
    for k in KEYS:
        rel = relerr(got[k].cpu(), ref[k])
        assert rel <= 5e-2, (k, rel)  # bf16 operands through ~70 chained contractions
    g = np.load(os.path.join(GOLD, "ssg_perturbed_288.npz"))  # outputs of the real reference
    assert relerr(got["protos"][:, ::4, ::4].cpu(), torch.from_numpy(g["protos_s"])) <= 5e-2
    assert relerr(got["grasp_coef_pred"][:, ::37].cpu(), torch.from_numpy(g["gcoef_s"])) <= 5e-2


def test_ssg_forward_full_size_init():
    """544 x 544, reference-style init, against the golden vectors of the real reference (fp32 mode)."""
    from crog_b200.model import SSG

    g = np.load(os.path.join(GOLD, "ssg_init_544.npz"))
    cfg = synth.ssg_cfg()
    model = SSG(cfg, precision="fp32")
    model.load_state_dict(synth.make_ssg_state_dict(cfg, 0, "init"), strict=True)
    model = model.cuda()
    rgb, depth = synth.make_ssg_inputs(1, 544)
    got = model({"rgb": rgb.cuda(), "depth": depth.cuda()})
    assert tuple(got["cls_pred"].shape) == (1, 18525, 32) and tuple(got["protos"].shape) == (1, 136, 136, 32)
    for k, gk in (("protos", "protos_s"), ("cls_pred", "cls_s"), ("box_pred", "box_s"), ("ins_coef_pred", "coef_s"), ("grasp_coef_pred", "gcoef_s")):
        sl = got[k][:, ::4, ::4] if k == "protos" else got[k][:, ::37]
        err = float((sl.cpu() - torch.from_numpy(g[gk])).abs().max())
        assert err <= 1e-3, (k, err)


def test_gaussian_bit_exact():
    from crog_b200.utils import grasp_eval as GE
    from oracle import ssg_forward as O

    rng = np.random.default_rng(1)
    maps = rng.random((3, 480, 640), dtype=np.float32)
    maps[rng.random(maps.shape) < 0.4] = 0.0
    got = GE.gaussian_batched(torch.from_numpy(maps).cuda(), 2.0).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], O.gaussian_f32(maps[i], 2.0))


@pytest.mark.parametrize("H,W", [(97, 133), (33, 7), (5, 300), (64, 128)])
def test_gaussian_forms_bit_identical(H, W):
    """One-kernel form (out of place), two-pass form (in place, through the intermediate buffer) and the oracle agree bit
    for bit on ragged sizes: maps narrower than the 17-tap window, widths that are no multiple of the 8-column groups,
    heights that are no multiple of the 32-row tiles."""
    from crog_b200.utils import grasp_eval as GE
    from oracle import ssg_forward as O

    rng = np.random.default_rng(H * 1000 + W)
    maps = rng.random((3, H, W), dtype=np.float32)
    d = torch.from_numpy(maps).cuda()
    fused = GE.gaussian_batched(d, 2.0)
    work = d.clone()
    GE.gaussian_batched(work, 2.0, out=work)  # in place: the two-pass kernels
    torch.cuda.synchronize()
    assert torch.equal(fused, work)
    for i in range(3):
        assert np.array_equal(fused[i].cpu().numpy(), O.gaussian_f32(maps[i], 2.0))


def test_fast_nms_wrapper_matches_oracle():
    from crog_b200.utils import grasp_eval as GE
    from oracle import ssg_forward as O

    cfg = synth.ssg_cfg(top_k=50, max_detections=20)
    g = torch.Generator().manual_seed(11)
    n = 700
    ctr = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(n, 2, generator=g) * 0.2 + 0.02
    box = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(0, 1)
    box[5] = torch.tensor([0.3, 0.3, 0.3, 0.3])  # degenerate: NaN IoU semantics
    cls = torch.rand(31, n, generator=g)
    cls[:, 100] = cls[:, 101]  # ties
    coef, gco = torch.randn(n, 32, generator=g), torch.randn(n, 4, 32, generator=g)
    want = O.fast_nms(cfg, box, cls, coef, gco)
    got = GE.fast_nms(cfg, box.cuda(), cls.cuda(), coef.cuda(), gco.cuda())
    assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
    assert torch.equal(got[2].cpu(), want[2]) and torch.equal(got[3].cpu(), want[3]) and torch.equal(got[4].cpu(), want[4])


def test_post_processing_matches_oracle_and_golden():
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail as T
    from oracle import ssg_forward as O

    cfg = synth.ssg_cfg()
    od = synth.make_ssg_output_dict(cfg, n_confident=8, seed=6)
    dd = {"ori_size": (480, 640)}
    ref = O.ssg_post_processing(cfg, od, dd, keep=True)
    got = GE.ssg_post_processing(cfg, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in od.items()}, dd)
    g = np.load(os.path.join(GOLD, "ssg_post_s6.npz"))  # written from the real reference's ssg_post_processing
    assert np.array_equal(got["cls"], ref["cls"]) and np.array_equal(got["cls"], g["cls"])  # detections, order, classes: exact
    assert np.abs(got["bboxes"] - ref["bboxes"]).max() <= 1e-3  # pixels; GPU expf vs CPU exp differ by an ulp
    n = len(ref["cls"])
    assert got["ins_masks"].shape == (n, 480, 640)
    # masks: fp32 dot products (32 terms, |proto| up to ~4) / sigmoid / bilinear in another summation order than the CPU: 1e-5 abs
    assert np.abs(got["ins_masks"] - ref["ins_masks"]).mean() <= 1e-5  # 0/1 masks may differ on the 0.5 iso-line only
    assert np.abs(got["grasp_masks"][2] - ref["grasp_masks"][2]).max() <= 1e-5
    assert np.abs(got["grasp_masks"][0] - ref["grasp_masks"][0]).max() <= 1e-5
    # decode: given the GPU's own smoothed maps, peaks and grasps must be what the oracle decodes from them, bit for bit
    qua, ang, wid = got["grasp_masks"]
    for i in range(n):
        # recover sin/cos at the peaks through the oracle's maps (angle column is compared through the angle map)
        peaks = T.peak_local_max(qua[i], 2, 0.4, 5)
        assert len(got["grasps_top5"][i]) == len(peaks)
        for r, (pr, pc) in zip(got["grasps_top5"][i], peaks):
            assert r[0] == float(pc) and r[1] == float(pr) and r[3] == 20
            assert r[2] == float(np.float64(wid[i][pr, pc]) * 100)
            assert abs(r[4] - float(np.float64(ang[i][pr, pc]) / np.pi * 180)) <= 1e-9
        assert got["grasps_top1"][i] == got["grasps_top5"][i][:1]
    # and the decoded grasp positions agree with the reference's on this fixture
    for a, b in zip(got["grasps_top5"], ref["grasps_top5"]):
        assert [(r[0], r[1]) for r in a] == [(r[0], r[1]) for r in b]


def test_post_processing_odd_original_size():
    """Mask assembly at an original size whose width is no multiple of 4 (scalar stores, ragged last column group) and whose
    height is no multiple of the 32-row blocks: same bars against the oracle as at 480 x 640."""
    from crog_b200.utils import grasp_eval as GE
    from oracle import ssg_forward as O

    cfg = synth.ssg_cfg()
    od = synth.make_ssg_output_dict(cfg, n_confident=5, seed=12)
    dd = {"ori_size": (241, 323)}
    ref = O.ssg_post_processing(cfg, od, dd, keep=True)
    got = GE.ssg_post_processing(cfg, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in od.items()}, dd)
    assert np.array_equal(got["cls"], ref["cls"])
    n = len(ref["cls"])
    assert n > 0 and got["ins_masks"].shape == (n, 241, 323)
    assert np.abs(got["ins_masks"] - ref["ins_masks"]).mean() <= 1e-5
    assert np.abs(got["grasp_masks"][2] - ref["grasp_masks"][2]).max() <= 1e-5
    assert np.abs(got["grasp_masks"][0] - ref["grasp_masks"][0]).max() <= 1e-5


def test_post_processing_batched_equals_per_image():
    """ssg_post_processing_batched (one sync, one map-major tensor for all instances of the batch) returns for every image
    exactly what the per-image drop-in ssg_post_processing returns (same kernels, different launch geometry)."""
    from crog_b200.utils import grasp_eval as GE

    cfg = synth.ssg_cfg()
    ods = [synth.make_ssg_output_dict(cfg, n_confident=n, seed=s) for n, s in ((8, 6), (0, 7), (3, 8), (11, 9))]
    batch = {k: torch.cat([od[k] for od in ods]).cuda() for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
    batch["anchors"] = ods[0]["anchors"]
    got = GE.ssg_post_processing_batched(cfg, batch, (480, 640))
    assert len(got) == len(ods)
    for od, g in zip(ods, got):
        ref = GE.ssg_post_processing(cfg, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in od.items()}, {"ori_size": (480, 640)})
        n = len(ref["cls"])
        assert g["n"] == n
        assert np.array_equal(g["cls"].cpu().numpy(), ref["cls"])
        assert np.array_equal(g["boxes"].cpu().numpy() * np.array([640, 640, 640, 640]), ref["bboxes"])
        hr = g["hr"].cpu().numpy()
        assert hr.shape == (5, n, 480, 640)
        if n == 0:
            continue
        assert np.array_equal(hr[0], ref["ins_masks"])
        assert np.array_equal(hr[1], ref["grasp_masks"][0]) and np.array_equal(hr[4], ref["grasp_masks"][2])
        npk, gr = g["n_peaks"].cpu().numpy(), g["grasps"].cpu().numpy()
        for i in range(n):
            rows = [[r[0], r[1], r[2], 20, r[4]] for r in gr[i, :npk[i]].tolist()]
            assert rows == ref["grasps_top5"][i]
