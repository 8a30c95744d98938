"""ctypes binding of libmila_b200_linear.so — the C-ABI declared in include/mila_b200_linear.h.

There is no fallback of any kind: if the shared library is missing or a call fails, this module
raises.  (PyTorch is used by the callers only for device memory and streams.)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# MILAB200_LIB selects another build of the same sources (the diagnostics build, `make -C mila_b200/csrc diag`)
LIB_PATH = Path(os.environ["MILAB200_LIB"]) if os.environ.get("MILAB200_LIB") else _PKG / "libmila_b200_linear.so"
_LIB: ctypes.CDLL | None = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64

# name -> argtypes ; every entry returns int (0 = ok)
SIGNATURES = {
    "milab200_quantize_fp8_per_channel": [c_p, c_p, c_p, c_i64, c_i64, c_p, c_p],
    "milab200_quantize_fp4_per_group": [c_p, c_p, c_p, c_i64, c_i64, c_i, c_p, c_p],
    "milab200_quantize_fp8_per_channel_device": [c_p, c_p, c_p, c_i64, c_i64, c_p],
    "milab200_quantize_fp4_per_group_device": [c_p, c_p, c_p, c_i64, c_i64, c_i, c_p],
    "milab200_matvec_decode_bf16_qfp8": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "milab200_matvec_decode_bf16_qfp4": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_w8a16_gemm": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_fp4a16_gemm": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_fp4a16_gemm_wmma": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_fp8_dequantize_to_bf16": [c_p, c_p, c_p, c_i, c_i, c_p],
    "milab200_fp4_dequantize_to_bf16": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_compute_fp8_weight_scale": [c_p, c_p, c_i64, c_p],
    "milab200_fp4_dequantize_to_fp8": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_quantize_bf16_to_fp8_per_token": [c_p, c_p, c_p, c_i, c_i, c_p],
    "milab200_fp8_apply_per_token_scales": [c_p, c_p, c_p, c_i, c_i, c_p],
    "milab200_add_bias_bf16": [c_p, c_p, c_i, c_i, c_p],
    "milab200_reserve_prefill": [c_i, c_i],
    "milab200_w8a16_gemm_glu": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_fp4a16_gemm_glu": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "milab200_geglu_forward_bf16": [c_p, c_p, c_i, c_i, c_p],
    "milab200_token_embedding_forward_bf16_qfp8": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_token_embedding_decode_bf16_qfp8": [c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "milab200_swiglu_forward_bf16": [c_p, c_p, c_i, c_i, c_p],
    "milab200_tp_create": [c_i, c_i, c_i, ctypes.POINTER(c_p)],
    "milab200_tp_export": [c_p, c_p],
    "milab200_tp_connect": [c_p, c_p],
    "milab200_tp_destroy": [c_p],
    "milab200_w8a16_gemm_rowparallel": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "milab200_fp4a16_gemm_rowparallel": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p],
    "milab200_rmsnorm_forward_bf16": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, ctypes.c_float, ctypes.c_float, c_p],
    "milab200_rmsnorm_w8a16_gemm": [c_p, c_p, c_p, c_p, c_p, ctypes.c_float, ctypes.c_float, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "milab200_rmsnorm_fp4a16_gemm": [c_p, c_p, c_p, c_p, c_p, ctypes.c_float, ctypes.c_float, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_rmsnorm_w8a16_gemm_glu": [c_p, c_p, c_p, c_p, c_p, c_p, ctypes.c_float, ctypes.c_float, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_rmsnorm_fp4a16_gemm_glu": [c_p, c_p, c_p, c_p, c_p, c_p, ctypes.c_float, ctypes.c_float, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "milab200_w4a16_gemm": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "milab200_w8a16_gemm_rowparallel_nccl": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "milab200_fp4a16_gemm_rowparallel_nccl": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p],
    "milab200_add_bias_f32": [c_p, c_p, c_i, c_i, c_p],
    "milab200_chain_create": [c_p, c_i, c_i, ctypes.POINTER(c_p)],
    "milab200_chain_forward": [c_p, c_p],
    "milab200_chain_destroy": [c_p],
    "milab200_chain_describe": [c_p, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i), ctypes.POINTER(c_i)],
    "milab200_chain_set_timeline": [c_p, c_p, ctypes.POINTER(c_i)],
}


class ChainLinear(ctypes.Structure):
    """milab200_chain_linear (include/mila_b200_linear.h)."""
    _fields_ = [("out_bf16", c_p), ("act_bf16", c_p), ("weight", c_p), ("scales", c_p), ("bias_bf16", c_p),
                ("in_features", c_i), ("out_features", c_i), ("group_size", c_i), ("glu", c_i), ("depends_on", c_i),
                ("tp_ctx", c_p)]
# exported but not returning a status
OTHER_SYMBOLS = ["milab200_abi_version", "milab200_error_string", "milab200_launch_count",
                 "milab200_reset_launch_count", "milab200_last_kernel", "milab200_init", "milab200_tp_handle_bytes",
                 "milab200_set_option", "milab200_note_weights_written"]


class MilaB200Error(RuntimeError):
    """std::runtime_error equivalent (CUDA failure / unsupported group size)."""


class InvalidArgument(ValueError):
    """std::invalid_argument equivalent (shape / argument errors)."""


class LogicError(RuntimeError):
    """std::logic_error equivalent (API misuse: backward on a quantized Linear, ...)."""


def build(verbose: bool = False) -> Path:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(_PKG / "csrc"), "-j8", "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmila_b200_linear.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise MilaB200Error(
                f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mila_b200/csrc`). There is no CPU fallback.")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, args in SIGNATURES.items():
            if os.environ.get("MILAB200_LIB") and not hasattr(L, name):
                continue          # another build of the sources (A/B against an older library): entries it lacks stay unbound
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = c_i
        L.milab200_abi_version.restype = c_i
        L.milab200_error_string.restype = ctypes.c_char_p
        L.milab200_error_string.argtypes = [c_i]
        L.milab200_launch_count.restype = ctypes.c_uint64
        L.milab200_reset_launch_count.restype = None
        L.milab200_last_kernel.restype = ctypes.c_char_p
        L.milab200_set_option.argtypes = [ctypes.c_char_p, c_i]
        L.milab200_set_option.restype = c_i
        L.milab200_note_weights_written.restype = c_i
        L.milab200_init.restype = c_i
        _LIB = L
    return _LIB


E_INVALID_ARGUMENT, E_UNSUPPORTED_GROUP, E_BAD_SHAPE, E_NO_DEVICE, E_NO_NCCL = -1, -2, -3, -4, -5


def check(rc: int, what: str) -> None:
    """Map a C-ABI return code onto the exception type the reference launcher would throw."""
    if rc == 0:
        return
    msg = f"{what}: {lib().milab200_error_string(rc).decode()} (code {rc})"
    if rc in (E_INVALID_ARGUMENT, E_BAD_SHAPE):
        raise InvalidArgument(msg)
    raise MilaB200Error(msg)


def set_option(name: str, value: int) -> None:
    """milab200_set_option: route selection (decode_tc, decode_mx4_max_m, decode_streamk, decode_presplit, decode_generic,
    prefill_tc, prefill_cta_group, prefill_act_planes, prefill_fp4_sum, prefill_glu, rmsnorm_fast_reduction, weights_written;
    include/mila_b200_linear.h lists them) — the programmatic form of the MILAB200_* environment switches."""
    check(lib().milab200_set_option(name.encode(), int(value)), f"set_option({name})")


def launch_count() -> int:
    return int(lib().milab200_launch_count())


def reset_launch_count() -> None:
    lib().milab200_reset_launch_count()


def last_kernel() -> str:
    return lib().milab200_last_kernel().decode()
