"""Helpers for the GPU parity tests: thin wrappers that call the C ABI directly."""
import ctypes as C

import torch

from crog_b200 import _lib as L


def pad_nhwc(x_nchw: torch.Tensor, dtype) -> torch.Tensor:
    """NCHW -> zero-haloed NHWC rows [B*(H+2)*(W+2), C]."""
    B, Cc, H, W = x_nchw.shape
    t = torch.zeros((B, H + 2, W + 2, Cc), device=x_nchw.device, dtype=dtype)
    t[:, 1:-1, 1:-1] = x_nchw.permute(0, 2, 3, 1).to(dtype)
    return t.reshape(-1, Cc).contiguous()


def compact_nhwc(x_nchw: torch.Tensor, dtype) -> torch.Tensor:
    B, Cc, H, W = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).to(dtype).reshape(-1, Cc).contiguous()


def unpad(rows: torch.Tensor, B, H, W) -> torch.Tensor:
    return rows.float().view(B, H + 2, W + 2, -1)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).contiguous()


def uncompact(rows: torch.Tensor, B, H, W) -> torch.Tensor:
    return rows.float().view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def conv_w(w: torch.Tensor, dtype) -> torch.Tensor:
    co = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(co, -1).to(dtype).contiguous()


def run_gemm(a, w, N, out, *, taps=1, cin=None, H=0, W=0, in_padded=False, out_padded=False, sample_rows=0, scale=None,
             bias=None, act=0, addmat=None, gate=None, scale2=None, bias2=None, residual=None, residual_relu=False,
             w_sample_stride=0, impl=0, a_col0=0, out_col0=0, tile_cfg=0, max_ctas=0, tap_mask=0, reverse=0):
    g = L.CrogGemm()
    es = a.element_size()
    g.a, g.a_rows, g.a_ld = a.data_ptr() + a_col0 * es, a.shape[0], a.shape[1]
    g.cin = cin if cin is not None else a.shape[1]
    g.taps, g.dtype, g.M, g.sample_rows, g.H, g.W = taps, L.dtype_code(a.dtype), a.shape[0], sample_rows, H, W
    g.in_padded, g.out_padded = int(in_padded), int(out_padded)
    g.w, g.N, g.w_sample_stride = w.data_ptr(), N, w_sample_stride
    g.scale = scale.data_ptr() if scale is not None else None
    g.bias = bias.data_ptr() if bias is not None else None
    if addmat is not None:
        g.addmat, g.addmat_rows = addmat.data_ptr(), addmat.shape[0]
    g.act = act
    if gate is not None:
        g.gate, g.scale2, g.bias2 = gate.data_ptr(), scale2.data_ptr(), bias2.data_ptr()
    if residual is not None:
        g.residual, g.res_ld, g.residual_relu = residual.data_ptr(), residual.stride(0), int(residual_relu)
    g.out, g.out_ld, g.out_dtype = out.data_ptr() + out_col0 * out.element_size(), out.shape[1], L.dtype_code(out.dtype)
    g.impl = impl
    g.tile_cfg = tile_cfg
    g.max_ctas, g.tap_mask, g.reverse = max_ctas, tap_mask, reverse
    L.check(L.lib().crog_gemm(C.byref(g), L.stream_ptr()))
    torch.cuda.synchronize()


def relerr(got: torch.Tensor, want: torch.Tensor) -> float:
    return float((got.float() - want.float()).norm() / (want.float().norm() + 1e-12))


def maxerr(got: torch.Tensor, want: torch.Tensor) -> float:
    return float((got.float() - want.float()).abs().max())
