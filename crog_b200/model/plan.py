"""Execution plan of the CROG forward on the crog_b200 kernels.

``ForwardPlan(sd, cfg, batch, precision)`` packs the reference state-dict once (BN folding,
K-major tap-ordered conv weights, positional terms folded into periodic bias matrices),
allocates every activation buffer for one batch size (NHWC rows, zero-haloed where a 3x3
convolution reads them; 180 GB of HBM means nothing is aliased or reused) and records the
kernel launches as a flat list that ``run()`` replays on the current stream, eagerly or
under CUDA-graph capture.  PyTorch is only the allocator / stream owner here; every op on
the per-batch path is a C-ABI call into libcrog_b200.so.

Layer semantics follow the reference line by line (citations at each stage).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from .. import _lib as L

BN_EPS = 1e-5
LN_EPS = 1e-5


class Act:
    """A [rows, ld] activation matrix of NHWC pixels (or plain rows when H == 0)."""

    def __init__(self, t: torch.Tensor, B: int, H: int, W: int, padded: bool, C_: int, col0: int = 0):
        self.t, self.B, self.H, self.W, self.padded, self.C, self.col0 = t, B, H, W, padded, C_, col0

    @property
    def ld(self) -> int:
        return self.t.shape[1]

    @property
    def rows(self) -> int:
        return self.t.shape[0]

    @property
    def sample_rows(self) -> int:
        if self.H == 0:
            return self.rows // max(self.B, 1)
        return (self.H + 2) * (self.W + 2) if self.padded else self.H * self.W

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.col0 * self.t.element_size()

    def cols(self, c0: int, c1: int) -> "Act":
        return Act(self.t, self.B, self.H, self.W, self.padded, c1 - c0, self.col0 + c0)

    def interior(self) -> torch.Tensor:
        """NCHW fp32 copy of the interior (debug / tests)."""
        t = self.t[:, self.col0:self.col0 + self.C].float()
        if self.H == 0:
            return t
        if self.padded:
            t = t.view(self.B, self.H + 2, self.W + 2, self.C)[:, 1:-1, 1:-1]
        else:
            t = t.view(self.B, self.H, self.W, self.C)
        return t.permute(0, 3, 1, 2).contiguous()


def _bn_fold(sd, p):
    s = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + BN_EPS)
    return s, sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * s


def _conv_w(w: torch.Tensor, cin_pad: int = 0, cout_pad: int = 0) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout(+pad), kh*kw*Cin(+pad)], tap-major then channel."""
    co, ci, kh, kw = w.shape
    w = w.permute(0, 2, 3, 1)
    if cin_pad > ci:
        w = F.pad(w, (0, cin_pad - ci))
    w = w.reshape(co, -1)
    if cout_pad > co:
        w = F.pad(w, (0, 0, 0, cout_pad - co))
    return w.contiguous()


def _pad_vec(v: torch.Tensor, n: int, fill: float) -> torch.Tensor:
    if v.numel() >= n:
        return v.contiguous()
    return torch.cat([v, torch.full((n - v.numel(),), fill, device=v.device, dtype=v.dtype)]).contiguous()


def _pos2d(d_model: int, H: int, W: int) -> torch.Tensor:
    """model/layers.py:216-241 -> [H*W, d_model] (input independent, built once)."""
    pe = torch.zeros(d_model, H, W)
    half = d_model // 2
    div = torch.exp(torch.arange(0.0, half, 2) * -(math.log(10000.0) / half))
    pw = torch.arange(0.0, W).unsqueeze(1) * div
    ph = torch.arange(0.0, H).unsqueeze(1) * div
    pe[0:half:2] = torch.sin(pw).t().unsqueeze(1).expand(-1, H, -1)
    pe[1:half:2] = torch.cos(pw).t().unsqueeze(1).expand(-1, H, -1)
    pe[half::2] = torch.sin(ph).t().unsqueeze(2).expand(-1, -1, W)
    pe[half + 1::2] = torch.cos(ph).t().unsqueeze(2).expand(-1, -1, W)
    return pe.reshape(d_model, H * W).t().contiguous()


def _pos1d(d_model: int, n: int) -> torch.Tensor:
    """model/layers.py:195-213 -> [n, d_model]."""
    pe = torch.zeros(n, d_model)
    position = torch.arange(0, n).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div)
    pe[:, 1::2] = torch.cos(position * div)
    return pe


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Kernels read parameters with 16-byte vector loads.  Tensors from torch's allocator are 256-byte aligned, but a
    DataParallel replica's parameters are views into one coalesced broadcast buffer and may start on any 4-byte boundary."""
    return t if t.data_ptr() % 16 == 0 else t.clone()


class ForwardPlan:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg, batch: int, precision: str = "bf16",
                 device: Optional[torch.device] = None, input_size: int = 416, gemm_impl: int = L.IMPL_AUTO,
                 keep: bool = False):
        assert precision in ("bf16", "fp32")
        self.lib = L.lib()
        self.cfg, self.B, self.precision = cfg, batch, precision
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.acode = L.dtype_code(self.adt)
        self.impl = gemm_impl
        self.ops: List[Callable[[int], None]] = []
        self.op_side: set = set()
        self.op_after: Dict[int, List[int]] = {}
        self.op_names: List[str] = []
        self.op_launches: List[int] = []  # kernel launches per recorded op (profiles/launch_list.py joins ncu rows on it)
        self.keep: Dict[str, Act] = {}
        self._keep_all = keep
        self._hold: List[object] = []  # keeps packed weights / descriptors alive
        self.n_launches = 0
        # bf16 tcgen05 plans merge the last convolution of a stage's first bottleneck with its downsample branch
        # (CROG_FUSE_DOWNSAMPLE=0: two GEMMs and an identity tensor, as in the fp32 mode)
        self.side_helpers = os.environ.get("CROG_SIDE_HELPERS", "1") != "0"
        self.fuse_downsample = (precision == "bf16" and gemm_impl != L.IMPL_SIMT and os.environ.get("CROG_FUSE_DOWNSAMPLE", "1") != "0")
        # bf16 tcgen05 plans store the two 32-channel stem tensors as pixel pairs (CROG_STEM_PAIRS=0: channel-padded rows)
        self.stem_pairs = (precision == "bf16" and gemm_impl != L.IMPL_SIMT and os.environ.get("CROG_STEM_PAIRS", "1") != "0")
        # SMs the text tower's persistent GEMMs may occupy while it runs beside the image front (0: no partition)
        self.text_sms = int(os.environ.get("CROG_TEXT_SMS", "16")) if precision == "bf16" and gemm_impl != L.IMPL_SIMT else 0
        # consecutive GEMMs walk their tiles in opposite directions: a consumer starts where its producer finished, so the
        # producer's most recent output is read from L2 instead of HBM (CROG_SNAKE=0: every GEMM walks upwards)
        self.snake = os.environ.get("CROG_SNAKE", "1") != "0"
        self._wdir = {}  # buffer address -> direction its producer GEMM walked
        self.front_end = 0  # ops [0, front_end) are the image front (stem, layer1, layer2) that overlaps the text tower
        self._side = None
        self._side2 = None
        self._sev = {}
        self._ev = None
        self.gemm_flops = 0
        self.gemm_alg_flops: Dict[str, int] = {}
        self.gemm_alg_bytes: Dict[str, int] = {}  # operands read once + result written once (no padding, no halo, no re-reads)
        self.gemm_ops: List[tuple] = []  # (name, descriptor, launch fn) of every GEMM, for the plan-time tile autotuner
        self.gemm_index: Dict[int, object] = {}  # op index -> descriptor
        self.tile_choice: Dict[str, tuple] = {}  # name -> (tile_cfg, us, heuristic us) once autotune() has run
        sd = {k: _aligned(v.detach().to(self.dev)) for k, v in sd.items()}
        self.sd = sd
        S = input_size
        assert S % 32 == 0
        self.S = S
        self.L_txt = cfg.word_len
        self.img = torch.zeros((batch, 3, S, S), device=self.dev, dtype=torch.float32)
        self.word = torch.zeros((batch, self.L_txt), device=self.dev, dtype=torch.int64)
        self._build()

    # ------------------------------------------------------------------ allocation helpers
    def new(self, H: int, W: int, Cc: int, padded: bool = False, dtype=None, rows: int = 0) -> Act:
        dtype = dtype or self.adt
        if H == 0:
            n = rows
        else:
            n = self.B * ((H + 2) * (W + 2) if padded else H * W)
        t = torch.zeros((n, Cc), device=self.dev, dtype=dtype)
        self._hold.append(t)  # ops capture raw pointers: the plan owns every buffer for its lifetime
        return Act(t, self.B, H, W, padded, Cc)

    def wt(self, t: torch.Tensor) -> torch.Tensor:
        t = _aligned(t.to(self.dev, self.adt).contiguous())
        self._hold.append(t)
        return t

    def f32(self, t: torch.Tensor) -> torch.Tensor:
        t = _aligned(t.to(self.dev, torch.float32).contiguous())
        self._hold.append(t)
        return t

    def _add(self, name: str, fn: Callable[[int], None], launches: int = 1, side: bool = False, after: Optional[List[int]] = None) -> int:
        """Record one op; returns its index.  side=True: a small memory-bound helper (pooling / layout copy) off the critical
        path - run() launches it on a second side stream, where it shares the SMs with the tensor-core kernels of the main
        stream (it needs no shared memory); after=[indices]: side ops whose results this op reads."""
        self.ops.append(fn)
        self.op_names.append(name)
        self.op_launches.append(launches)
        self.n_launches += launches
        idx = len(self.ops) - 1
        if side:
            self.op_side.add(idx)
        if after:
            self.op_after[idx] = list(after)
        return idx

    # ------------------------------------------------------------------ op recorders
    def gemm(self, name: str, a: Act, w: torch.Tensor, N: int, out: Act, taps: int = 1, scale=None, bias=None,
             act: int = L.ACT_NONE, residual: Optional[Act] = None, residual_relu: bool = False, addmat=None,
             gate=None, scale2=None, bias2=None, w_sample_stride: int = 0, cin: Optional[int] = None,
             alg_n: Optional[int] = None, alg_cin: Optional[int] = None, out_sample_rows: int = 0, out_row0: int = 0,
             row_stats_out: Optional[torch.Tensor] = None, row_stats_in: Optional[torch.Tensor] = None, row_stats_width: int = 0,
             a2: Optional[Act] = None, after: Optional[List[int]] = None, tap_mask: int = 0,
             alg_flops: Optional[int] = None, alg_bytes: Optional[int] = None):
        g = L.CrogGemm()
        g.tap_mask = tap_mask
        if getattr(self, "snake", False):
            g.reverse = 1 - self._wdir.get(a.ptr, 0)
            self._wdir[out.ptr] = g.reverse
        cin = cin if cin is not None else a.C
        cin2 = a2.C if a2 is not None else 0
        assert w.shape[-1] == taps * cin + cin2, (name, tuple(w.shape), taps, cin, cin2)
        if a2 is not None:  # second operand contracted after the first (two summed 1x1 convolutions as one GEMM)
            assert taps == 1 and a2.rows == a.rows and a2.padded == a.padded and (a2.H, a2.W) == (a.H, a.W), name
            g.a2, g.a2_ld, g.cin2 = a2.ptr, a2.ld, cin2
        g.a, g.a_rows, g.a_ld, g.cin, g.taps, g.dtype = a.ptr, a.rows, a.ld, cin, taps, self.acode
        g.M, g.sample_rows, g.H, g.W = a.rows, a.sample_rows, a.H, a.W
        g.in_padded, g.out_padded = int(a.padded), int(out.padded)
        g.w, g.N, g.w_sample_stride = w.data_ptr(), N, w_sample_stride
        g.scale = scale.data_ptr() if scale is not None else None
        g.bias = bias.data_ptr() if bias is not None else None
        if addmat is not None:
            assert addmat.shape[1] == N and addmat.dtype == torch.float32
            g.addmat, g.addmat_rows = addmat.data_ptr(), addmat.shape[0]
        g.act = act
        if gate is not None:
            g.gate, g.scale2, g.bias2 = gate.data_ptr(), scale2.data_ptr(), bias2.data_ptr()
        if residual is not None:
            assert residual.t.dtype == out.t.dtype and residual.padded == out.padded
            g.residual, g.res_ld, g.residual_relu = residual.ptr, residual.ld, int(residual_relu)
        g.out, g.out_ld, g.out_dtype = out.ptr + out_row0 * out.ld * out.t.element_size(), out.ld, L.dtype_code(out.t.dtype)
        g.impl = self.impl
        g.out_sample_rows = out_sample_rows
        if row_stats_out is not None:  # folded LayerNorm, producer side: (sum, sum^2) per row and 64-column chunk
            assert row_stats_out.dtype == torch.float32 and row_stats_out.shape == (a.rows, N // 64, 2)
            g.row_stats_out, g.row_stats_chunks, g.row_stats_width, g.row_stats_eps = row_stats_out.data_ptr(), N // 64, N, LN_EPS
        if row_stats_in is not None:   # consumer side: scale = s, bias = c (see _decoder)
            assert row_stats_in.dtype == torch.float32 and row_stats_in.shape[0] == a.rows and scale is not None and bias is not None
            g.row_stats_in, g.row_stats_chunks, g.row_stats_width, g.row_stats_eps = (row_stats_in.data_ptr(), row_stats_in.shape[1],
                                                                                   row_stats_width, LN_EPS)
        self._hold.extend([row_stats_out, row_stats_in])
        if a.H > 0 and out_sample_rows == 0:
            assert out.H == a.H and out.W == a.W, name
        self._hold.extend([g, w, scale, bias, addmat, gate, scale2, bias2])
        lib = self.lib
        ref = C.byref(g)
        idx = self._add(name, lambda s: L.check(lib.crog_gemm(ref, s)), after=after)
        self.gemm_ops.append((name, g, self.ops[-1]))
        self.gemm_index[idx] = g
        rows_eff = a.B * a.H * a.W if a.H > 0 else a.rows
        self.gemm_flops += 2 * rows_eff * N * (taps * cin + cin2)
        # algorithmic FLOPs of the reference op (no channel padding, no halo rows) for roofline accounting
        self.gemm_alg_flops[name] = alg_flops if alg_flops is not None else 2 * rows_eff * (alg_n or N) * (taps * (alg_cin or cin) + cin2)
        esz = a.t.element_size()
        w_copies = (a.rows // a.sample_rows) if w_sample_stride > 0 else 1
        self.gemm_alg_bytes[name] = (rows_eff * ((alg_cin or cin) + cin2) * esz + w_copies * (alg_n or N) * (taps * (alg_cin or cin) + cin2) * esz
                                     + rows_eff * (alg_n or N) * out.t.element_size()
                                     + (rows_eff * N * residual.t.element_size() if residual is not None else 0))
        if alg_bytes is not None:
            self.gemm_alg_bytes[name] = alg_bytes
        if self._keep_all:
            self.keep[name] = out

    def resample(self, name: str, src: Act, dst: Act, mode: int, side: bool = False) -> int:
        lib, code = self.lib, L.dtype_code(src.t.dtype)
        assert src.t.dtype == dst.t.dtype
        a = (src.ptr, src.ld, int(src.padded), dst.ptr, dst.ld, int(dst.padded), src.B, src.H, src.W, src.C, mode, code)
        idx = self._add(name, lambda s: L.check(lib.crog_resample(*a, s)), side=side and self.side_helpers)
        if self._keep_all:
            self.keep[name] = dst
        return idx

    def layernorm(self, name: str, x: Act, p: str, out: Act, residual: Optional[Act] = None):
        lib = self.lib
        g, b = self.f32(self.sd[p + ".weight"]), self.f32(self.sd[p + ".bias"])
        assert x.ld == x.C and out.ld == out.C, "layernorm works on dense rows"
        a = (x.ptr, L.dtype_code(x.t.dtype), g.data_ptr(), b.data_ptr(), residual.ptr if residual else None, out.ptr,
             L.dtype_code(out.t.dtype), x.rows, x.C, LN_EPS)
        self._add(name, lambda s: L.check(lib.crog_layernorm(*a, s)))
        if self._keep_all:
            self.keep[name] = out

    def layernorm_chain(self, name: str, x: Act, p1: str, stream_: Act, p2: str, out: Act):
        """stream_ += LN(x; p1); out = LN(stream_; p2) in one pass (csrc/elementwise.cu: layernorm_chain_kernel)."""
        lib = self.lib
        g1, b1 = self.f32(self.sd[p1 + ".weight"]), self.f32(self.sd[p1 + ".bias"])
        g2, b2 = self.f32(self.sd[p2 + ".weight"]), self.f32(self.sd[p2 + ".bias"])
        assert x.ld == x.C and out.ld == out.C and stream_.ld == stream_.C and stream_.t.dtype == torch.float32
        a = (x.ptr, L.dtype_code(x.t.dtype), g1.data_ptr(), b1.data_ptr(), stream_.ptr, stream_.ptr, g2.data_ptr(), b2.data_ptr(),
             out.ptr, L.dtype_code(out.t.dtype), x.rows, x.C, LN_EPS)
        self._add(name, lambda s: L.check(lib.crog_layernorm_chain(*a, s)))
        if self._keep_all:
            self.keep[name] = out

    def attention(self, name: str, q: Act, k: Act, v: Act, o: Act, heads: int, Tq: int, Tk: int, causal: bool = False,
                  pad_word: Optional[torch.Tensor] = None):
        lib = self.lib
        a = (q.ptr, q.ld, k.ptr, k.ld, v.ptr, v.ld, o.ptr, o.ld, self.B, heads, Tq, Tk, 0.125, int(causal),
             pad_word.data_ptr() if pad_word is not None else None, self.acode)
        self._add(name, lambda s: L.check(lib.crog_attention(*a, s)))
        self.gemm_flops += 4 * self.B * heads * Tq * Tk * 64
        if self._keep_all:
            self.keep[name] = o

    # ------------------------------------------------------------------ the network
    def _build(self):
        sd, cfg, B, S = self.sd, self.cfg, self.B, self.S
        lib = self.lib
        RELU = L.ACT_RELU
        # ---- stem (clip.py:165-184, 208-213): conv1 on CUDA cores, conv2/conv3 as implicit GEMM with the
        # 32 channels zero-padded to one 64-channel K chunk
        v = "backbone.visual"
        H1 = S // 2
        sc1, bi1 = _bn_fold(sd, v + ".bn1")
        w1, sc1, bi1 = self.f32(sd[v + ".conv1.weight"]), self.f32(sc1), self.f32(bi1)
        sc2, bi2 = _bn_fold(sd, v + ".bn2")
        sc3, bi3 = _bn_fold(sd, v + ".bn3")
        s3 = self.new(H1, H1, 64)
        esz = 2 if self.adt == torch.bfloat16 else 4
        npx = B * H1 * H1
        if self.stem_pairs and H1 % 2 == 0:
            # PIXEL-PAIR layout for the two 32-channel stem tensors: a row holds two horizontally adjacent pixels (2 x 32
            # channels = one full 64-channel K chunk, no padding channels), on a zero-haloed [H1+2, H1/2+2] grid.  A 3x3
            # convolution over pixels is a 3x3 convolution over pairs whose weights place w[o, i, ky, dx] at pair offset ksx
            # / input parity q / output parity p with dx = 2 ksx + q - p (zero where |dx| > 1):
            #   conv2 (32 -> 32): one GEMM, N = 64 = 2 parities x 32, K = 9 x 64 - half the rows and half the executed FLOPs
            #     of the channel-padded form (which contracted 64 x 64 per tap for a 32 x 32 convolution);
            #   conv3 (32 -> 64): one GEMM per output parity (N = 64) writing its half of the [pixels/2, 128] view of the
            #     ordinary NHWC output; each parity reaches only two of the three pair offsets, the third tap is masked out
            #     (6 of 9 taps contracted).
            Wp = H1 // 2
            s1 = self.new(H1, Wp, 64, padded=True)
            a = (self.img.data_ptr(), B, S, S, w1.data_ptr(), sc1.data_ptr(), bi1.data_ptr(), 32, s1.ptr, s1.ld, self.acode, 1)
            self._add("stem.conv1", lambda s: L.check(lib.crog_stem_conv1(*a, s)))
            w2, w3 = sd[v + ".conv2.weight"].float(), sd[v + ".conv3.weight"].float()

            def pair_weights(w, parities):
                co, ci = w.shape[0], w.shape[1]
                out = torch.zeros((len(parities) * co, 3, 3, 2 * ci), device=w.device)
                mask = 0
                for pi, p_ in enumerate(parities):
                    for q in (0, 1):
                        for ksx in (-1, 0, 1):
                            dx = 2 * ksx + q - p_
                            if abs(dx) <= 1:
                                out[pi * co:(pi + 1) * co, :, ksx + 1, q * ci:(q + 1) * ci] = w[:, :, :, dx + 1].permute(0, 2, 1)
                                for ky in range(3):
                                    mask |= 1 << (ky * 3 + ksx + 1)
                return out.reshape(len(parities) * co, -1).contiguous(), mask

            w2p, _ = pair_weights(w2, (0, 1))
            s2 = self.new(H1, Wp, 64, padded=True)
            self.gemm("stem.conv2", s1, self.wt(w2p), 64, s2, taps=9, scale=self.f32(torch.cat([sc2, sc2])),
                      bias=self.f32(torch.cat([bi2, bi2])), act=RELU, alg_flops=2 * npx * 32 * 9 * 32,
                      alg_bytes=npx * 32 * esz * 2 + 32 * 288 * esz)
            s3v = Act(s3.t.view(-1, 128), B, H1, Wp, False, 128)
            for p_ in (0, 1):
                w3p, mask = pair_weights(w3, (p_,))
                self.gemm(f"stem.conv3.p{p_}", s2, self.wt(w3p), 64, s3v.cols(64 * p_, 64 * p_ + 64), taps=9, scale=self.f32(sc3),
                          bias=self.f32(bi3), act=RELU, tap_mask=mask, alg_flops=npx * 64 * 9 * 32,
                          alg_bytes=(npx * 32 * esz + npx * 64 * esz + 64 * 288 * esz) // 2)
        else:
            # channel-padded form (fp32 / CUDA-core plans): 32 channels zero-padded to one 64-channel K chunk
            s1 = self.new(H1, H1, 64, padded=True)
            a = (self.img.data_ptr(), B, S, S, w1.data_ptr(), sc1.data_ptr(), bi1.data_ptr(), 32, s1.ptr, s1.ld, self.acode, 0)
            self._add("stem.conv1", lambda s: L.check(lib.crog_stem_conv1(*a, s)))
            s2 = self.new(H1, H1, 64, padded=True)
            self.gemm("stem.conv2", s1, self.wt(_conv_w(sd[v + ".conv2.weight"], 64, 64)), 64, s2, taps=9,
                      scale=self.f32(_pad_vec(sc2, 64, 1.0)), bias=self.f32(_pad_vec(bi2, 64, 0.0)), act=RELU, alg_n=32, alg_cin=32)
            self.gemm("stem.conv3", s2, self.wt(_conv_w(sd[v + ".conv3.weight"], 64)), 64, s3, taps=9,
                      scale=self.f32(sc3), bias=self.f32(bi3), act=RELU, alg_cin=32)
        x = self.new(H1 // 2, H1 // 2, 64)
        self.resample("stem.avgpool", s3, x, L.RS_AVGPOOL2)
        self.keep["stem"] = x
        # ---- residual stages (clip.py:44-57, 187-203)
        feats = []
        inpl = 64
        for li, nb in enumerate((3, 4, 6, 3), start=1):
            planes = 64 * 2 ** (li - 1)
            for bi_ in range(nb):
                x = self._bottleneck(f"{v}.layer{li}.{bi_}", x, inpl, planes, 2 if (li > 1 and bi_ == 0) else 1)
                inpl = planes * 4
            feats.append(x)
            self.keep[f"layer{li}"] = x
            # zero-haloed copies of C3 / C4 for the 3x3 convolutions of the neck: produced on the helper stream as soon
            # as the stage is done, under the tensor-core work of the following stages
            if li == 2:
                self.front_end = len(self.ops)
                self._c3p = self.new(x.H, x.W, x.C, padded=True)
                self._c3p_idx = self.resample("neck.c3_pad", x, self._c3p, L.RS_COPY, side=True)
            if li == 3:
                self._c4p = self.new(x.H, x.W, x.C, padded=True)
                self._c4p_idx = self.resample("neck.c4_pad", x, self._c4p, L.RS_COPY, side=True)
        c3, c4, x4 = feats[1], feats[2], feats[3]
        c5 = self._attnpool(x4)
        self.keep.update(c3=c3, c4=c4, c5=c5)
        t_begin = len(self.ops)
        wordfeat, state32, state_a = self._text()
        # everything that only depends on the sentence state rides on the text branch too: the FPN text gate
        # (layers.py:376) and the projector's text-generated kernel folded with vis.4 (layers.py:91-93)
        self._text_gate(state_a)
        self._dynamic_weights(state32)
        self.text_range = (t_begin, len(self.ops))  # independent of the image tower: may run on a forked stream
        fq = self._neck(c3, c4, c5)
        if not cfg.use_contrastive:
            self.keep["fq_neck"] = fq  # (with a decoder this buffer is its in-place fp32 residual stream)
        if cfg.use_contrastive:
            fq = self._decoder(fq, wordfeat)
            self.keep["fq_dec"] = fq
        self._projector(fq)

    def _bottleneck(self, p: str, x: Act, inpl: int, planes: int, stride: int) -> Act:
        sd, RELU = self.sd, L.ACT_RELU
        H, W = x.H, x.W
        # the pooled input of the downsample branch is only needed by the last GEMM of the block: it is produced on the
        # helper stream while conv1 / conv2 run
        has_ds = (p + ".downsample.0.weight") in sd
        xi, pool_idx = x, None
        if has_ds and stride > 1:
            xi = self.new(H // 2, W // 2, inpl)
            pool_idx = self.resample(p + ".downsample.pool", x, xi, L.RS_AVGPOOL2, side=True)
        dep = [pool_idx] if pool_idx is not None else None
        sc, bi = _bn_fold(sd, p + ".bn1")
        t1 = self.new(H, W, planes, padded=True)
        self.gemm(p + ".conv1", x, self.wt(_conv_w(sd[p + ".conv1.weight"])), planes, t1, scale=self.f32(sc),
                  bias=self.f32(bi), act=RELU)
        sc, bi = _bn_fold(sd, p + ".bn2")
        t2 = self.new(H, W, planes)
        self.gemm(p + ".conv2", t1, self.wt(_conv_w(sd[p + ".conv2.weight"])), planes, t2, taps=9, scale=self.f32(sc),
                  bias=self.f32(bi), act=RELU)
        if stride > 1:
            t2p = self.new(H // 2, W // 2, planes)
            self.resample(p + ".avgpool", t2, t2p, L.RS_AVGPOOL2)
            t2 = t2p
        idt = x
        out = self.new(t2.H, t2.W, planes * 4)
        if has_ds and self.fuse_downsample:
            # out = relu(bn3(conv3(t2)) + bn_d(conv_d(xi))) as ONE contraction over [t2 | xi] (K = planes + inpl) with both
            # BatchNorm scales folded into the bf16 weights and the two shifts summed: the identity tensor of the first block of
            # every stage (up to 354 MB written and read back) and one epilogue pass disappear
            s3, b3 = _bn_fold(sd, p + ".bn3")
            s_d, b_d = _bn_fold(sd, p + ".downsample.1")
            wcat = torch.cat([_conv_w(sd[p + ".conv3.weight"]).float() * s3[:, None],
                              _conv_w(sd[p + ".downsample.0.weight"]).float() * s_d[:, None]], 1)
            self.gemm(p + ".conv3+downsample", t2, self.wt(wcat), planes * 4, out, bias=self.f32(b3 + b_d), act=RELU, a2=xi, after=dep)
            return out
        if has_ds:
            sc, bi = _bn_fold(sd, p + ".downsample.1")
            idt = self.new(xi.H, xi.W, planes * 4)
            self.gemm(p + ".downsample", xi, self.wt(_conv_w(sd[p + ".downsample.0.weight"])), planes * 4, idt,
                      scale=self.f32(sc), bias=self.f32(bi), after=dep)
        sc, bi = _bn_fold(sd, p + ".bn3")
        self.gemm(p + ".conv3", t2, self.wt(_conv_w(sd[p + ".conv3.weight"])), planes * 4, out, scale=self.f32(sc),
                  bias=self.f32(bi), residual=idt, residual_relu=True)
        return out

    def _attnpool(self, x4: Act) -> Act:
        """clip.py:110-144 (modified AttentionPool2d: no mean token, `connect` residual, ReLU)."""
        sd, a = self.sd, "backbone.visual.attnpool"
        H, W, E = x4.H, x4.W, x4.C
        T = H * W
        fuse = self.fuse_downsample  # same trick as the bottlenecks: c_proj(att) + connect(x4) as one GEMM over [att | x4]
        sc, bi = _bn_fold(sd, a + ".connect.1")
        res = None
        if not fuse:
            res = self.new(H, W, 1024)
            self.gemm(a + ".connect", x4, self.wt(_conv_w(sd[a + ".connect.0.weight"])), 1024, res, scale=self.f32(sc),
                      bias=self.f32(bi))
        pe = sd[a + ".positional_embedding"].float()
        g = int(round(math.sqrt(pe.shape[0] - 1)))
        pe = pe[1:].reshape(1, g, g, E).permute(0, 3, 1, 2)
        pe = F.interpolate(pe, size=(H, W), mode="bicubic", align_corners=False)  # input independent (clip.py:101-104)
        pe = pe.flatten(2)[0].t().contiguous()  # [T, E]
        wqkv = torch.cat([sd[a + ".q_proj.weight"], sd[a + ".k_proj.weight"], sd[a + ".v_proj.weight"]]).float()
        bqkv = torch.cat([sd[a + ".q_proj.bias"], sd[a + ".k_proj.bias"], sd[a + ".v_proj.bias"]]).float()
        addmat = self.f32(pe @ wqkv.t() + bqkv)  # (x + pos) W^T + b = x W^T + (pos W^T + b)
        qkv = self.new(0, 0, 3 * E, rows=self.B * T)
        xin = Act(x4.t, self.B, H, W, False, E)
        self.gemm(a + ".qkv", xin, self.wt(wqkv), 3 * E, Act(qkv.t, self.B, H, W, False, 3 * E), addmat=addmat)
        att = self.new(0, 0, E, rows=self.B * T)
        self.attention(a + ".attn", qkv.cols(0, E), qkv.cols(E, 2 * E), qkv.cols(2 * E, 3 * E), att, E // 64, T, T)
        c5 = self.new(H, W, 1024)
        if fuse:
            wcat = torch.cat([sd[a + ".c_proj.weight"].float(), _conv_w(sd[a + ".connect.0.weight"]).float() * sc[:, None]], 1)
            self.gemm(a + ".c_proj+connect", Act(att.t, self.B, H, W, False, E), self.wt(wcat), 1024, c5,
                      bias=self.f32(sd[a + ".c_proj.bias"].float() + bi), act=L.ACT_RELU, a2=x4)
        else:
            self.gemm(a + ".c_proj", Act(att.t, self.B, H, W, False, E), self.wt(sd[a + ".c_proj.weight"].float()), 1024, c5,
                      bias=self.f32(sd[a + ".c_proj.bias"]), residual=res, residual_relu=True)
        return c5

    def _text(self):
        """clip.py:439-456 + ResidualAttentionBlock clip.py:239-265.  fp32 residual stream."""
        sd, lib, B, Lt = self.sd, self.lib, self.B, self.L_txt
        D = sd["backbone.ln_final.weight"].shape[0]
        heads = D // 64
        rows = B * Lt
        x = self.new(0, 0, D, dtype=torch.float32, rows=rows)
        emb, pos = self.f32(sd["backbone.token_embedding.weight"]), self.f32(sd["backbone.positional_embedding"])
        a = (self.word.data_ptr(), emb.data_ptr(), pos.data_ptr(), x.ptr, B, Lt, D, emb.shape[0])
        self._add("text.embed", lambda s: L.check(lib.crog_embed_tokens(*a, s)))
        h = self.new(0, 0, D, rows=rows)
        qkv = self.new(0, 0, 3 * D, rows=rows)
        att = self.new(0, 0, D, rows=rows)
        ff = self.new(0, 0, 4 * D, rows=rows)
        i = 0
        while f"backbone.transformer.resblocks.{i}.ln_1.weight" in sd:
            p = f"backbone.transformer.resblocks.{i}"
            self.layernorm(p + ".ln_1", x, p + ".ln_1", h)
            self.gemm(p + ".attn.in_proj", h, self.wt(sd[p + ".attn.in_proj_weight"]), 3 * D, qkv,
                      bias=self.f32(sd[p + ".attn.in_proj_bias"]))
            self.attention(p + ".attn", qkv.cols(0, D), qkv.cols(D, 2 * D), qkv.cols(2 * D, 3 * D), att, heads, Lt, Lt, causal=True)
            self.gemm(p + ".attn.out_proj", att, self.wt(sd[p + ".attn.out_proj.weight"]), D, x,
                      bias=self.f32(sd[p + ".attn.out_proj.bias"]), residual=x)
            self.layernorm(p + ".ln_2", x, p + ".ln_2", h)
            self.gemm(p + ".mlp.c_fc", h, self.wt(sd[p + ".mlp.c_fc.weight"]), 4 * D, ff,
                      bias=self.f32(sd[p + ".mlp.c_fc.bias"]), act=L.ACT_QUICKGELU)
            self.gemm(p + ".mlp.c_proj", ff, self.wt(sd[p + ".mlp.c_proj.weight"]), D, x,
                      bias=self.f32(sd[p + ".mlp.c_proj.bias"]), residual=x)
            i += 1
        wordfeat = self.new(0, 0, D, rows=rows)
        self.layernorm("text.ln_final", x, "backbone.ln_final", wordfeat)
        eot = self.new(0, 0, D, rows=B)
        a2 = (self.word.data_ptr(), wordfeat.ptr, self.acode, eot.ptr, self.acode, B, Lt, D)
        self._add("text.eot", lambda s: L.check(lib.crog_gather_eot(*a2, s)))
        E = sd["backbone.text_projection"].shape[1]
        state32 = self.new(0, 0, E, dtype=torch.float32, rows=B)
        self.gemm("text.projection", eot, self.wt(sd["backbone.text_projection"].float().t()), E, state32)
        state_a = state32
        if self.adt != torch.float32:
            state_a = self.new(0, 0, E, rows=B)
            a3 = (state32.ptr, L.F32, state_a.ptr, self.acode, B * E)
            self._add("text.state_cast", lambda s: L.check(lib.crog_cast(*a3, s)))
        self.keep.update(word=wordfeat, state=state32)
        return wordfeat, state32, state_a

    def _cbr(self, name: str, p: str, a: Act, out: Act, taps: int, **kw):
        """conv(no bias) + BN + ReLU (layers.py:8-11)."""
        sc, bi = _bn_fold(self.sd, p + ".1")
        wsrc = self.sd[p + ".0.weight"]
        self.gemm(name, a, self.wt(_conv_w(wsrc)), wsrc.shape[0], out, taps=taps, scale=self.f32(sc), bias=self.f32(bi),
                  act=L.ACT_RELU, **kw)

    def _text_gate(self, state_a: Act):
        """layers.py:376: state' = ReLU(BN1d(W . state)), the per-(sample, channel) gate of the FPN."""
        sd = self.sd
        fo = list(self.cfg.fpn_out)
        sc, bi = _bn_fold(sd, "neck.txt_proj.1")
        self.gate = self.new(0, 0, fo[2], dtype=torch.float32, rows=self.B)
        self.gemm("neck.txt_proj", state_a, self.wt(sd["neck.txt_proj.0.weight"]), fo[2], self.gate, scale=self.f32(sc),
                  bias=self.f32(bi), act=L.ACT_RELU)

    def _neck(self, c3: Act, c4: Act, c5: Act) -> Act:
        """layers.py:371-398."""
        sd, B = self.sd, self.B
        fo = list(self.cfg.fpn_out)
        gate = self.gate
        s2, b2 = _bn_fold(sd, "neck.norm_layer.0")
        f5 = self.new(c5.H, c5.W, fo[2], padded=True)
        self._cbr("neck.f1_v_proj", "neck.f1_v_proj", c5, f5, 1, gate=gate.t, scale2=self.f32(s2), bias2=self.f32(b2))
        H4, W4 = c4.H, c4.W
        c4p = self._c4p
        cat4 = self.new(H4, W4, fo[1] + fo[2])
        self._cbr("neck.f2_v_proj", "neck.f2_v_proj", c4p, cat4.cols(0, fo[1]), 9, after=[self._c4p_idx])
        self.resample("neck.f5_up", f5, cat4.cols(fo[1], fo[1] + fo[2]), L.RS_BILINEAR2)
        cat3 = self.new(H4, W4, fo[0] + fo[1])
        f4 = cat3.cols(fo[0], fo[0] + fo[1])
        self._cbr("neck.f2_cat", "neck.f2_cat", cat4, f4, 1)
        c3p = self._c3p
        t3 = self.new(c3.H, c3.W, fo[0])
        self._cbr("neck.f3_v_proj", "neck.f3_v_proj", c3p, t3, 9, after=[self._c3p_idx])
        self.resample("neck.f3_pool", t3, cat3.cols(0, fo[0]), L.RS_AVGPOOL2)
        f3p = self.new(H4, W4, fo[1], padded=True)
        self._cbr("neck.f3_cat", "neck.f3_cat", cat3, f3p, 1)
        f4p = self.new(H4, W4, fo[1], padded=True)
        self.resample("neck.f4_pad", f4, f4p, L.RS_COPY)
        catq = self.new(H4, W4, 3 * fo[1])
        q5 = self.new(c5.H, c5.W, fo[1])
        self._cbr("neck.f4_proj5", "neck.f4_proj5", f5, q5, 9)
        self.resample("neck.fq5_up", q5, catq.cols(2 * fo[1], 3 * fo[1]), L.RS_BILINEAR2)
        self._cbr("neck.f4_proj4", "neck.f4_proj4", f4p, catq.cols(fo[1], 2 * fo[1]), 9)
        self._cbr("neck.f4_proj3", "neck.f4_proj3", f3p, catq.cols(0, fo[1]), 9)
        ag = self.new(H4, W4, fo[1], padded=True)
        self._cbr("neck.aggr", "neck.aggr", catq, ag, 1)
        # CoordConv (layers.py:19-44): the two coordinate channels are input independent, so their
        # convolution is a per-pixel term added before BN
        wcc = sd["neck.coordconv.0.conv1.0.weight"].float()
        ys = torch.linspace(-1, 1, H4, device=self.dev).view(1, 1, H4, 1).expand(1, 1, H4, W4)
        xs = torch.linspace(-1, 1, W4, device=self.dev).view(1, 1, 1, W4).expand(1, 1, H4, W4)
        cmap = F.conv2d(torch.cat([xs, ys], 1), wcc[:, fo[1]:], padding=1)[0]  # [Cout, H, W]
        cmap = self.f32(cmap.flatten(1).t())  # [H*W, Cout]
        sc, bi = _bn_fold(sd, "neck.coordconv.0.conv1.1")
        cc = self.new(H4, W4, fo[1], padded=True)
        self.gemm("neck.coordconv.0", ag, self.wt(_conv_w(wcc[:, :fo[1]])), fo[1], cc, taps=9, scale=self.f32(sc),
                  bias=self.f32(bi), act=L.ACT_RELU, addmat=cmap)
        out_dtype = torch.float32 if self.cfg.use_contrastive else self.adt
        fq = self.new(H4, W4, fo[1], dtype=out_dtype)
        self._cbr("neck.coordconv.1", "neck.coordconv.1", cc, fq, 9)
        return fq

    def _decoder(self, vis: Act, wordfeat: Act) -> Act:
        """layers.py:243-277, 313-339.  `vis` is the fp32 residual stream [B*HW, D]."""
        sd, B, cfg = self.sd, self.B, self.cfg
        D, Hh, Ww, Lt, heads = vis.C, vis.H, vis.W, self.L_txt, cfg.num_head
        T = Hh * Ww
        vp = _pos2d(D, Hh, Ww).to(self.dev)
        tp = _pos1d(D, Lt).to(self.dev)
        rows = B * T
        v2 = self.new(Hh, Ww, D)
        qkv = self.new(Hh, Ww, 3 * D)
        att = self.new(Hh, Ww, D)
        tmp = self.new(Hh, Ww, D)
        qc = self.new(Hh, Ww, D)
        kv = self.new(0, 0, 2 * D, rows=B * Lt)
        ff = self.new(Hh, Ww, cfg.dim_ffn)
        # bf16 tcgen05 plans fold the FFN LayerNorm into ffn.4 (CROG_FFN_LN_FOLD=0: keep the separate LayerNorm pass)
        fold_ln = (self.precision == "bf16" and self.impl != L.IMPL_SIMT and cfg.dim_ffn % 64 == 0 and
                   os.environ.get("CROG_FFN_LN_FOLD", "1") != "0")
        ff2 = None if fold_ln else self.new(Hh, Ww, cfg.dim_ffn)
        ffn_stats = torch.zeros((rows, cfg.dim_ffn // 64, 2), device=self.dev, dtype=torch.float32) if fold_ln else None
        self._hold.append(ffn_stats)
        i = 0
        while f"decoder.layers.{i}.norm1.weight" in sd:
            p = f"decoder.layers.{i}"
            w_in, b_in = sd[p + ".self_attn.in_proj_weight"].float(), sd[p + ".self_attn.in_proj_bias"].float()
            add = torch.cat([vp @ w_in[:2 * D].t(), torch.zeros(T, D, device=self.dev)], 1) + b_in
            self.layernorm(p + ".norm1", vis, p + ".norm1", v2)
            self.gemm(p + ".self_attn.in_proj", v2, self.wt(w_in), 3 * D, qkv, addmat=self.f32(add))
            self.attention(p + ".self_attn", qkv.cols(0, D), qkv.cols(D, 2 * D), qkv.cols(2 * D, 3 * D), att, heads, T, T)
            self.gemm(p + ".self_attn.out_proj", att, self.wt(sd[p + ".self_attn.out_proj.weight"]), D, tmp,
                      bias=self.f32(sd[p + ".self_attn.out_proj.bias"]))
            # vis += LN(self-attention output); v2 = norm2(vis): chained in one pass (layers.py:318-322)
            self.layernorm_chain(p + ".self_attn_norm+norm2", tmp, p + ".self_attn_norm", vis, p + ".norm2", v2)
            w_in, b_in = sd[p + ".multihead_attn.in_proj_weight"].float(), sd[p + ".multihead_attn.in_proj_bias"].float()
            self.gemm(p + ".cross.q", v2, self.wt(w_in[:D]), D, qc, addmat=self.f32(vp @ w_in[:D].t() + b_in[:D]))
            addkv = torch.cat([tp @ w_in[D:2 * D].t(), torch.zeros(Lt, D, device=self.dev)], 1) + b_in[D:]
            self.gemm(p + ".cross.kv", wordfeat, self.wt(w_in[D:]), 2 * D, kv, addmat=self.f32(addkv))
            self.attention(p + ".cross_attn", qc, kv.cols(0, D), kv.cols(D, 2 * D), att, heads, T, Lt, pad_word=self.word)
            self.gemm(p + ".cross.out_proj", att, self.wt(sd[p + ".multihead_attn.out_proj.weight"]), D, tmp,
                      bias=self.f32(sd[p + ".multihead_attn.out_proj.bias"]))
            self.layernorm_chain(p + ".cross_attn_norm+norm3", tmp, p + ".cross_attn_norm", vis, p + ".norm3", v2)
            if fold_ln:
                # Linear -> ReLU -> LayerNorm(2048) -> Linear with the LayerNorm folded into the second contraction:
                # LN(h) W^T = rstd (h W'^T) - mu rstd s + c,  W' = W gamma (per input column), s = row sums of the bf16 W',
                # c = W beta + b.  ffn.0 also emits the per-row (sum h, sum h^2) partials; the 354 MB LayerNorm pass is gone.
                w4, gam, bet = sd[p + ".ffn.4.weight"].float(), sd[p + ".ffn.3.weight"].float(), sd[p + ".ffn.3.bias"].float()
                w4g = self.wt(w4 * gam[None, :])
                s_vec = self.f32(w4g.float().sum(1))
                c_vec = self.f32(w4 @ bet + sd[p + ".ffn.4.bias"].float())
                self.gemm(p + ".ffn.0", v2, self.wt(sd[p + ".ffn.0.weight"]), cfg.dim_ffn, ff, bias=self.f32(sd[p + ".ffn.0.bias"]),
                          act=L.ACT_RELU, row_stats_out=ffn_stats)
                self.gemm(p + ".ffn.4", ff, w4g, D, vis, scale=s_vec, bias=c_vec, residual=vis, row_stats_in=ffn_stats,
                          row_stats_width=cfg.dim_ffn)
            else:
                self.gemm(p + ".ffn.0", v2, self.wt(sd[p + ".ffn.0.weight"]), cfg.dim_ffn, ff, bias=self.f32(sd[p + ".ffn.0.bias"]),
                          act=L.ACT_RELU)
                self.layernorm(p + ".ffn.3", ff, p + ".ffn.3", ff2)
                self.gemm(p + ".ffn.4", ff2, self.wt(sd[p + ".ffn.4.weight"]), D, vis, bias=self.f32(sd[p + ".ffn.4.bias"]),
                          residual=vis)
            i += 1
        out = self.new(Hh, Ww, D)
        self.layernorm("decoder.norm", vis, "decoder.norm", out)
        return out

    def _dynamic_weights(self, state32: Act):
        """layers.py:91-93 folded with vis.4 (layers.py:58,70-77): per-sample [ZR, CP] kernel of the head convolution."""
        sd, B, lib = self.sd, self.B, self.lib
        Cc = sd["proj.vis.3.0.weight"].shape[0]
        NH = sd["proj.vis.4.weight"].shape[0] // Cc
        ZR, CP = ((9 * NH + 15) // 16) * 16, Cc + 64  # Z columns (9 taps x NH heads, padded), feature channels + ones chunk
        wfold = torch.zeros((B * ZR, CP), device=self.dev, dtype=self.adt)
        scratch = torch.zeros((B, 9 * Cc + 1), device=self.dev, dtype=torch.float32)
        tw, tb = self.f32(sd["proj.txt.weight"]), self.f32(sd["proj.txt.bias"])
        vw = self.f32(sd["proj.vis.4.weight"].reshape(NH * Cc, Cc))
        vb = self.f32(sd["proj.vis.4.bias"])
        a = (state32.ptr, L.F32, tw.data_ptr(), tb.data_ptr(), vw.data_ptr(), vb.data_ptr(), scratch.data_ptr(),
             wfold.data_ptr(), self.acode, B, state32.C, Cc, NH, ZR, CP)
        self._hold.extend([wfold, scratch])
        self._add("proj.dynw_fold", lambda s: L.check(lib.crog_dynw_fold(*a, s)), launches=2)
        self.wfold, self.NH, self._proj_dims = wfold, NH, (Cc, ZR, CP)

    def _projector(self, fq: Act):
        """layers.py:64-132 (MultiTaskProjector) / :152-173 (Projector).  vis.4 (1x1, 256->256*NH) is folded
        into the text-generated 3x3 kernel, so the heads are one per-sample 3x3 convolution."""
        sd, B, lib = self.sd, self.B, self.lib
        NH, wfold = self.NH, self.wfold
        Cc, ZR, CP = self._proj_dims
        H1, W1 = fq.H * 2, fq.W * 2
        u1 = self.new(H1, W1, fq.C, padded=True)
        self.resample("proj.up1", fq, u1, L.RS_BILINEAR2)
        p1 = self.new(H1, W1, sd["proj.vis.1.0.weight"].shape[0])
        self._cbr("proj.vis.1", "proj.vis.1", u1, p1, 9)
        H2, W2 = H1 * 2, W1 * 2
        u2 = self.new(H2, W2, p1.C, padded=True)
        self.resample("proj.up2", p1, u2, L.RS_BILINEAR2)
        feat = self.new(H2, W2, CP, padded=True)
        feat.t.view(B, H2 + 2, W2 + 2, CP)[:, 1:-1, 1:-1, Cc] = 1.0  # constant-one channel carrying the biases
        self._cbr("proj.vis.3", "proj.vis.3", u2, feat.cols(0, Cc), 9)
        # per-sample 1x1 GEMM: Z[p, h*9+tap] = feat[p, :] . wfold[b, h*9+tap, :]  (features are read once, not nine times)
        z = self.new(H2, W2, ZR, padded=True, dtype=torch.float32)
        self.gemm("proj.dynconv", feat, wfold, ZR, z, taps=1, w_sample_stride=ZR * CP, cin=CP, alg_n=9 * NH, alg_cin=Cc)
        self.out = torch.zeros((NH, B, 1, H2, W2), device=self.dev, dtype=torch.float32)
        a2 = (z.ptr, z.ld, self.out.data_ptr(), B, H2, W2, NH)
        self._add("proj.gather", lambda s: L.check(lib.crog_dynconv_gather(*a2, s)))

    # ------------------------------------------------------------------ plan-time tile autotuner
    def autotune(self, reps: int = 4, min_gain: float = 0.03) -> Dict[str, tuple]:
        """Pick the tcgen05 tile configuration of every GEMM of this plan by timing the applicable CROG_TILE_* ones on
        the plan's own buffers (CUDA events on the current stream).  All configurations accumulate the k-blocks in the
        same order and share the epilogue, so the choice changes speed only, never a bit of the result
        (tests/test_gpu_kernels.py::test_tile_cfgs_bit_identical).  A forced configuration replaces the built-in
        heuristic only when it is more than `min_gain` faster.  Launch-identical layers (same shape, layout and
        epilogue) are timed once."""
        if self.precision != "bf16" or self.impl == L.IMPL_SIMT:
            return {}
        lib = self.lib
        s = L.stream_ptr()
        # realistic operands: one forward on random inputs (zero buffers would flatter every configuration)
        self._random_inputs()
        self.run(stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def clock(fn) -> float:
            fn(s)  # first launch of an instantiation sets its attributes; also warms the operands' L2 lines
            best = float("inf")
            for _ in range(2):
                e0.record()
                for _ in range(reps):
                    fn(s)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
            return best

        memo: Dict[tuple, tuple] = {}
        for name, g, fn in self.gemm_ops:
            key = (g.M, g.N, g.cin, g.taps, g.a_ld, g.out_ld, g.res_ld, g.H, g.W, g.in_padded, g.out_padded, g.out_dtype,
                   g.w_sample_stride, g.out_sample_rows, g.act, bool(g.residual), bool(g.addmat), bool(g.gate), bool(g.scale))
            if key not in memo:
                g.tile_cfg = L.TILE_AUTO
                base = clock(fn)
                best_cfg, best_us = L.TILE_AUTO, base
                for cfg_id in range(1, min(L.TILE_COUNT, int(os.environ.get("CROG_AUTOTUNE_MAX_CFG", L.TILE_COUNT - 1)) + 1)):
                    g.tile_cfg = cfg_id
                    if lib.crog_gemm(C.byref(g), s) != 0:
                        continue  # configuration does not apply to this shape
                    us = clock(fn)
                    if us < best_us:
                        best_cfg, best_us = cfg_id, us
                if best_us > base * (1.0 - min_gain):
                    best_cfg, best_us = L.TILE_AUTO, base
                memo[key] = (best_cfg, best_us, base)
            g.tile_cfg = memo[key][0]
            self.tile_choice[name] = memo[key]
        torch.cuda.synchronize()
        self._zero_inputs()
        return self.tile_choice

    def _random_inputs(self):
        gen = torch.Generator(device="cpu").manual_seed(1234)
        self.img.copy_(torch.randn(self.img.shape, generator=gen))
        w = torch.zeros(self.word.shape, dtype=torch.int64)
        w[:, 0], w[:, 1:5], w[:, 5] = 49406, 1000, 49407
        self.word.copy_(w)

    def _zero_inputs(self):
        self.img.zero_()
        self.word.zero_()

    # ------------------------------------------------------------------ execution
    def run(self, stream: Optional[int] = None, fork_text: bool = True):
        """Replay the recorded launches on the current stream.  The text encoder does not depend on the image tower
        until the neck, so it is forked onto a side stream (under CUDA-graph capture this becomes a parallel branch)."""
        if stream is not None or not fork_text or os.environ.get("CROG_NO_FORK"):
            self._partition(False)
            s = stream if stream is not None else L.stream_ptr()
            for fn in self.ops:
                fn(s)
            return
        self._partition(True)
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
            self._side2 = torch.cuda.Stream(device=self.dev)
            self._ev = (torch.cuda.Event(), torch.cuda.Event())
            self._sev = {i: (torch.cuda.Event(), torch.cuda.Event()) for i in self.op_side}
        t0, t1 = self.text_range
        self._ev[0].record(main)
        self._side.wait_event(self._ev[0])
        if not os.environ.get("CROG_DEBUG_SKIP_TEXT"):  # timing experiments only: the image path alone (results are garbage)
            for fn in self.ops[t0:t1]:
                fn(self._side.cuda_stream)
        self._ev[1].record(self._side)

        def on_main(lo, hi):
            for i in range(lo, hi):
                if i in self.op_side:  # helper stream: starts once the main stream has reached this point
                    self._sev[i][0].record(main)
                    self._side2.wait_event(self._sev[i][0])
                    self.ops[i](self._side2.cuda_stream)
                    self._sev[i][1].record(self._side2)
                    continue
                for d in self.op_after.get(i, ()):
                    if d in self.op_side:
                        main.wait_event(self._sev[d][1])
                self.ops[i](main.cuda_stream)

        on_main(0, t0)
        main.wait_event(self._ev[1])
        on_main(t1, len(self.ops))

    def _partition(self, on: bool):
        """SM budgets of the two concurrent branches.  Persistent GEMM CTAs take most of an SM's shared memory, so the text
        tower's and the image tower's cannot share an SM: left alone, the two branches mostly serialise.  While the text
        tower runs (the memory-bound image front: stem, layer1, layer2) its GEMMs get `text_sms` SMs and the front's the
        rest; a serial replay (per-op timing, autotune, debugging) gives every GEMM the whole GPU.  The descriptors are
        read at launch time, so a captured graph keeps the budgets it was captured with."""
        t0, t1 = self.text_range
        n_sm = torch.cuda.get_device_properties(self.dev).multi_processor_count
        for i, g in self.gemm_index.items():
            cap = 0
            if on and self.text_sms > 0 and t1 > t0:
                if t0 <= i < t1:
                    cap = self.text_sms
                elif i < self.front_end:
                    cap = n_sm - self.text_sms
            g.max_ctas = cap

    def run_debug(self):
        """Run op by op with a device sync after each, naming the op that faults."""
        for name, fn in zip(self.op_names, self.ops):
            fn(L.stream_ptr())
            try:
                torch.cuda.synchronize()
            except Exception as e:  # pragma: no cover
                raise L.CrogError(f"kernel fault in op '{name}': {e}") from e
