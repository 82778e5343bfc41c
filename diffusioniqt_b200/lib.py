"""ctypes binding of libdiqt_b200.so (the C ABI declared in include/diqt.h).

There is no fallback: if the shared library is missing or a call fails, this module raises.
Pointers are raw device addresses (`tensor.data_ptr()`), the stream is torch's current CUDA
stream, so every call is capturable in a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIQT_LIB_PATH") or os.path.join(_HERE, "libdiqt_b200.so")   # the variable selects an A/B build variant

F32, BF16 = 0, 1
CONV_K3, CONV_K1, CONV_DOWN, CONV_UP = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_ZM = 0, 1, 2, 3
CONV_FLAG_NO_CTA_PAIR = 1   # include/diqt.h DIQT_CONV_FLAG_NO_CTA_PAIR
ABI_VERSION = 1


class DiqtError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mode", "dtype", "impl", "n", "d0", "d1", "d2", "c_in", "ld_in", "c_out", "ld_out", "flags")]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes ; every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    "diqt_abi_version": [],
    "diqt_last_error": [],
    "diqt_launch_count": [],
    "diqt_conv_resolved_impl": [C.POINTER(ConvDesc), C.POINTER(C.c_int)],
    "diqt_conv_packed_bytes": [C.POINTER(ConvDesc), C.POINTER(C.c_size_t)],
    "diqt_conv_pack": [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp],
    "diqt_conv_plan_create": [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, C.POINTER(_vp)],
    "diqt_conv_plan_set_stats": [_vp, _vp, C.POINTER(C.c_int)],
    "diqt_conv_plan_workspace_bytes": [_vp, C.POINTER(C.c_size_t)],
    "diqt_conv_plan_set_workspace": [_vp, _vp, C.c_size_t],
    "diqt_conv_plan_destroy": [_vp],
    "diqt_conv_run": [_vp, _vp],
    "diqt_channel_stats": [_vp, _i, _i, _i64, _i, _i, _i, _vp, _i, _i, _vp],
    "diqt_gn_finalize": [_vp, _i, _i, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "diqt_affine_mish": [_vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _vp, _i, _i, _i, _vp],
    "diqt_se_gate": [_vp, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "diqt_scale_residual": [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _i, _vp, _i, _i, _vp],
    "diqt_stats_groups": [_i, _i, C.POINTER(C.c_int)],
    "diqt_conv_plan_set_stats_g": [_vp, _vp, _vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "diqt_channel_stats_g": [_vp, _i, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp],
    "diqt_conv_gn_fusable": [C.POINTER(ConvDesc)],
    "diqt_conv_plan_set_gn": [_vp, _vp, _i, _i64, _i, _f, _vp, _vp],
    "diqt_conv_plan_set_film": [_vp, _vp, _i, _vp, _i],
    "diqt_conv_plan_set_gn_affine": [_vp, _vp, _vp],
    "diqt_gn_mish_g": [_vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _i, _i, _f, _vp, _vp, _vp, _i, _vp, _i, _i, _vp],
    "diqt_scale_residual_g": [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp],
    "diqt_scale_copy": [_vp, _i, _vp, _i, _i, _i64, _i, _f, _vp],
    "diqt_chan_layernorm": [_vp, _i, _vp, _i, _i, _i64, _i, _vp, _vp, _f, _i, _vp, _i, _vp, _i, _i, _i, _vp],
    "diqt_rows_combine": [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _vp],
    "diqt_dw_patchify": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp],
    "diqt_dw_conv3": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "diqt_upsample_trilinear": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "diqt_linear_attention_chunks": [_i, C.POINTER(C.c_int)],
    "diqt_linear_attention": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp],
    "diqt_softmax_attention": [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _i, _vp],
    "diqt_attn_tc_supported": [_i, _i, _i, _i, _i, _i],
    "diqt_attn_tc_workspace_bytes": [_i, _i, C.POINTER(C.c_size_t)],
    "diqt_attn_tc_plan_create": [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f, _i, _vp, C.POINTER(_vp)],
    "diqt_attn_tc_plan_destroy": [_vp],
    "diqt_attn_tc_run": [_vp, _vp],
    "diqt_linattn_tc_supported": [_i, _i, _i, _i, _i],
    "diqt_linattn_tc_workspace_bytes": [_i, _i, C.POINTER(C.c_size_t)],
    "diqt_linattn_tc_plan_create": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _f, _i, _vp, C.POINTER(_vp)],
    "diqt_linattn_tc_plan_destroy": [_vp],
    "diqt_linattn_tc_run": [_vp, _vp],
    "diqt_bwd_reduce": [_vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _vp, _i, _i, _vp, _vp],
    "diqt_bwd_apply": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "diqt_gn_bwd_finalize": [_vp, _i, _vp, _i, _i, _i64, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "diqt_se_bwd": [_vp, _i, _vp, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "diqt_conv_wgrad_workspace_bytes": [_i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_size_t)],
    "diqt_conv_wgrad_resolved_impl": [_i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_int)],
    "diqt_conv_wgrad": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "diqt_loss_grad": [_vp, _vp, _i, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _vp],
    "diqt_adam_step": [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _f, _vp],
    "diqt_init_conv_pack": [_vp, _i, _i, _vp, _vp],
    "diqt_init_conv_k": [C.POINTER(_vp), C.POINTER(_i64), _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "diqt_init_conv_tc_supported": [_i, _i, _i, _i],
    "diqt_init_conv_tc_blocks": [_i, _i, _i, C.POINTER(C.c_int)],
    "diqt_init_conv_tc": [C.POINTER(_vp), C.POINTER(_i64), _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "diqt_init_im2col": [C.POINTER(_vp), C.POINTER(_i64), _i, _vp, _i, _i, _i, _i, _vp],
    "diqt_init_conv": [C.POINTER(_vp), C.POINTER(_i64), _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "diqt_final_conv": [_vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "diqt_ddpm_update": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "diqt_edm_prepare": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "diqt_final_conv_edm": [_vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "diqt_edm_update": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "diqt_clamp": [_vp, _i64, _f, _f, _vp],
    "diqt_fourier_features": [_vp, _i, _vp, _i, _vp, _vp],
    "diqt_linear": [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "diqt_advance_step": [_vp, _vp],
    "diqt_gather_patches": [_vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp],
    "diqt_stitch_patches": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _f, _vp],
}
_RESTYPES = {"diqt_last_error": C.c_char_p, "diqt_launch_count": C.c_uint64, "diqt_conv_plan_destroy": None, "diqt_attn_tc_plan_destroy": None,
             "diqt_linattn_tc_plan_destroy": None}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the kernel library once; raise loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DiqtError(
            f"{LIB_PATH} not found: the CUDA kernel library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or diffusioniqt_b200/csrc/build.sh). "
            "There is no CPU or PyTorch fallback for the sampling path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.diqt_abi_version() != ABI_VERSION:
        raise DiqtError(f"libdiqt_b200.so ABI {lib.diqt_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().diqt_last_error()
        raise DiqtError(f"{what or 'diqt call'} failed ({rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().diqt_launch_count())


def ptr(t) -> int:
    """Device address of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
