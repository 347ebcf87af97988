"""ctypes binding of libpdm.so (C ABI: include/pdm.h).

There is no CPU fallback: if the shared library is missing this module raises, and every compute
entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDM_LIB") or os.path.join(HERE, "libpdm.so")  # PDM_LIB: development override

PREC_BF16 = 0
PREC_FP32 = 1
FWD_GROUND_TRUTH = 1  # PDM_FWD_GROUND_TRUTH
PLAN_STRIDE = 16
ABI_VERSION = 2
METHOD_CODES = {"fast": 0, "singlestep": 1, "multistep": 2}
SKIP_CODES = {"time_uniform": 0, "logSNR": 1, "t2": 2}


class PdmConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "img_size", "patch_size", "in_chans", "embed_dim", "depth", "num_heads", "mlp_ratio", "clip_dim",
        "num_clip_token", "num_panoptic_class", "enable_panoptic", "separate")]


class PdmVaeConfig(C.Structure):
    _fields_ = [("ch", C.c_int32), ("num_levels", C.c_int32), ("ch_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32),
                ("z_channels", C.c_int32), ("embed_dim", C.c_int32), ("out_ch", C.c_int32), ("scale_factor", C.c_float)]


# symbol -> (restype, argtypes); mirrors include/pdm.h one to one
_P = C.c_void_p
SIGNATURES = {
    "pdm_create": (C.c_int, [C.POINTER(PdmConfig), C.POINTER(_P)]),
    "pdm_destroy": (C.c_int, [_P]),
    "pdm_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, _P]),
    "pdm_finalize_params": (C.c_int, [_P, _P]),
    "pdm_workspace_bytes": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "pdm_nnet_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "pdm_nnet_forward_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "pdm_cfg_update": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_float), C.c_float,
                                 C.c_int64, C.c_int64, _P]),
    "pdm_solver_plan": (C.c_int, [C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_float, C.c_int32, C.c_float, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]),
    "pdm_multistep_update": (C.c_int, [_P] * 14 + [C.POINTER(C.c_float), C.c_float, C.c_int64, C.c_int64, _P]),
    "pdm_sample": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int32, _P, _P, _P, _P, C.c_float, _P, _P, C.c_int32,
                             C.c_int32, C.c_int32, _P]),
    "pdm_bits2int": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "pdm_int2bits": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "pdm_debug_linear": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.POINTER(C.c_float), _P]),
    "pdm_debug_attention": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), _P]),
    "pdm_debug_layernorm": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), _P]),
    "pdm_debug_ln_chain": (C.c_int, [_P] * 10 + [C.c_int32] * 6 + [C.POINTER(C.c_float), _P]),
    "pdm_vae_create": (C.c_int, [C.POINTER(PdmVaeConfig), C.POINTER(_P)]),
    "pdm_vae_destroy": (C.c_int, [_P]),
    "pdm_vae_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, _P]),
    "pdm_vae_finalize_params": (C.c_int, [_P, _P]),
    "pdm_vae_decode": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "pdm_vae_workspace_bytes": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "pdm_last_error": (C.c_char_p, []),
    "pdm_abi_version": (C.c_int, []),
    "pdm_launch_count": (C.c_int64, []),
    "pdm_set_profiling": (C.c_int, [_P, C.c_int32]),
    "pdm_get_profile": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libpdm.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found. Build it with `python -m panopticdiffusionmodels_b200.build` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if handle.pdm_abi_version() != ABI_VERSION:
            raise RuntimeError("libpdm.so ABI version mismatch; rebuild the library")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().pdm_last_error()
        raise RuntimeError("libpdm: " + (msg.decode() if msg else "unknown error"))


def ptr(t) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libpdm has no CPU path: tensor must live on a CUDA device")
    if not t.is_contiguous():
        raise RuntimeError("libpdm needs contiguous tensors")
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().pdm_launch_count())
