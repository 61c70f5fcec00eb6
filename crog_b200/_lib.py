"""ctypes binding of libcrog_b200.so (include/crog_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every
entry point raises.  The library is built in-tree by ``crog_b200.build`` (or
``__graft_entry__.build()``); importing this module never compiles anything.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CROG_B200_SO") or os.path.join(_HERE, "lib", "libcrog_b200.so")  # override: A/B of two builds

ABI_VERSION = 6  # crog_abi_version() of the library these signatures were written for
F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_QUICKGELU, ACT_TANH = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
# CROG_TILE_* of include/crog_b200.h (CrogGemm.tile_cfg)
(TILE_AUTO, TILE_128x128, TILE_128x256, TILE_128x256_E8, TILE_PAIR_256x256, TILE_PAIR_256x256_E8, TILE_PAIR_256x128,
 TILE_128x64, TILE_CONV3, TILE_128x128_S3, TILE_128x128_E12, TILE_BAND_PAIR_256x256, TILE_BAND_PAIR_256x256_E8,
 TILE_BAND_PAIR_256x128, TILE_BAND_128x128, TILE_BAND_128x256, TILE_CONV3_E12, TILE_CONV3_PAIR, TILE_CONV3_DUAL,
 TILE_COUNT) = range(20)
RS_COPY, RS_AVGPOOL2, RS_BILINEAR2, RS_SUBSAMPLE2, RS_BILINEAR2_AC = 0, 1, 2, 3, 4


class CrogError(RuntimeError):
    pass


class CrogGemm(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int64), ("a_ld", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32),
        ("dtype", C.c_int32), ("M", C.c_int32), ("sample_rows", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("in_padded", C.c_int32), ("out_padded", C.c_int32), ("w", C.c_void_p), ("N", C.c_int32),
        ("w_sample_stride", C.c_int64), ("scale", C.c_void_p), ("bias", C.c_void_p), ("addmat", C.c_void_p),
        ("addmat_rows", C.c_int32), ("act", C.c_int32), ("gate", C.c_void_p), ("scale2", C.c_void_p),
        ("bias2", C.c_void_p), ("residual", C.c_void_p), ("res_ld", C.c_int32), ("residual_relu", C.c_int32),
        ("out", C.c_void_p), ("out_ld", C.c_int32), ("out_dtype", C.c_int32), ("impl", C.c_int32),
        ("out_sample_rows", C.c_int32), ("tile_cfg", C.c_int32),
        ("a2", C.c_void_p), ("a2_ld", C.c_int32), ("cin2", C.c_int32),
        ("row_stats_out", C.c_void_p), ("row_stats_in", C.c_void_p), ("row_stats_chunks", C.c_int32),
        ("row_stats_width", C.c_int32), ("row_stats_eps", C.c_float),
        ("max_ctas", C.c_int32), ("tap_mask", C.c_int32), ("reverse", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/crog_b200.h declares
_P, _I, _L, _F, _U = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32
SIGNATURES = {
    "crog_last_error": (C.c_char_p, []),
    "crog_abi_version": (C.c_int, []),
    "crog_check_device": (C.c_int, []),
    "crog_gemm": (C.c_int, [C.POINTER(CrogGemm), _P]),
    "crog_resample": (C.c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "crog_stem_conv1": (C.c_int, [_P, _I, _I, _I, _P, _P, _P, _I, _P, _I, _I, _I, _P]),
    "crog_layernorm": (C.c_int, [_P, _I, _P, _P, _P, _P, _I, _L, _I, _F, _P]),
    "crog_layernorm_chain": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _F, _P]),
    "crog_embed_tokens": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "crog_gather_eot": (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _P]),
    "crog_attention": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _P, _I, _P]),
    "crog_dynw_fold": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "crog_dynconv_gather": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "crog_cast": (C.c_int, [_P, _I, _P, _I, _L, _P]),
    "crog_split_heads": (C.c_int, [_P, _I, _P, _L, _I, _P]),
    "crog_sigmoid_bicubic": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _U, _P]),
    "crog_detect_workspace_bytes": (C.c_int64, [_I, _I, _I, _I]),
    "crog_detect_grasps": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P]),
    "crog_angle_map": (C.c_int, [_P, _P, _P, _L, _P]),
    "crog_jaccard": (C.c_int, [_P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P]),
    "crog_stem7_patches": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "crog_maxpool3s2": (C.c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "crog_patches3": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "crog_ssg_heads": (C.c_int, [_P, _I, _L, _I, _I, _P, _P, _P]),
    "crog_ssg_nms_workspace_bytes": (C.c_int64, [_I, _I]),
    "crog_ssg_fast_nms": (C.c_int, [_P, _P, _P, _I, _I, _F, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "crog_ssg_detect": (C.c_int, [_P, _P, _P, _I, _I, _F, _F, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "crog_ssg_masks": (C.c_int, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "crog_ssg_detect_batched": (C.c_int, [_P, _P, _P, _I, _I, _I, _F, _F, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "crog_ssg_masks_batched": (C.c_int, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "crog_gaussian": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    "crog_warp_affine_cubic_f32": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _F, _P]),
    "crog_preprocess_workspace_bytes": (C.c_int64, []),
    "crog_preprocess_u8": (C.c_int, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "crog_mask_iou": (C.c_int, [_P, _P, _I, _L, _F, _P, _P]),
}

_lib = None
_device_ok = set()


def load() -> C.CDLL:
    """Load the shared library (no device needed) and bind every declared symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CrogError(f"{SO_PATH} is missing: run `python -m crog_b200.build` (or __graft_entry__.build()). "
                            "crog_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(SO_PATH)
        lib.crog_abi_version.restype = C.c_int
        if lib.crog_abi_version() != ABI_VERSION:  # a stale build would be called with shifted arguments
            raise CrogError(f"{SO_PATH} has ABI version {lib.crog_abi_version()}, this package binds version {ABI_VERSION}: "
                            "rebuild with `python -m crog_b200.build`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib() -> C.CDLL:
    """Library handle for launching kernels: also checks the current device is a B200."""
    import torch

    l = load()
    if not torch.cuda.is_available():
        raise CrogError("crog_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _device_ok:
        check(l.crog_check_device())
        _device_ok.add(dev)
    return l


def check(rc: int) -> None:
    if rc != 0:
        msg = load().crog_last_error().decode("utf-8", "replace")
        raise CrogError(f"crog_b200 error {rc}: {msg}")


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def dtype_code(t) -> int:
    import torch

    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise CrogError(f"unsupported dtype {t}")
