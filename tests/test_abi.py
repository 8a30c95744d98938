"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares; argument errors
come back as codes (no compute is attempted without a GPU); the host mirror keeps the reference's
error conventions."""
import ctypes
import re
from pathlib import Path

import pytest

from mila_b200 import _lib
from mila_b200.linear import Linear, LinearConfig, PerChannelFp8, PerGroupFp4

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names |= set(re.findall(r"\b(milab200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported():
    L = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"
    # and the binding table covers every status-returning entry
    for s in syms:
        assert s in _lib.SIGNATURES or s in _lib.OTHER_SYMBOLS, s


def test_every_reference_launcher_has_a_slot():
    # symbols a replacement must provide
    for ref in ("quantize_fp8_per_channel", "quantize_fp4_per_group", "matvec_decode_bf16_qfp8",
                "matvec_decode_bf16_qfp4", "w8a16_gemm", "fp4a16_gemm", "fp4a16_gemm_wmma",
                "compute_fp8_weight_scale", "fp4_dequantize_to_fp8", "quantize_bf16_to_fp8_per_token",
                "fp8_apply_per_token_scales", "fp8_dequantize_to_bf16", "fp4_dequantize_to_bf16"):
        assert "milab200_" + ref in _lib.SIGNATURES
    wrappers = (ROOT / "include" / "mila_b200" / "Kernels").rglob("*.cuh")
    text = "".join(p.read_text() for p in wrappers)
    for ref in ("cuda_quantize_fp8_per_channel", "cuda_quantize_fp4_per_group", "cuda_matvec_decode_bf16_qfp8",
                "cuda_matvec_decode_bf16_qfp4", "cuda_w8a16_gemm", "cuda_fp4a16_gemm", "cuda_fp4a16_gemm_wmma"):
        assert re.search(r"\b" + ref + r"\s*\(", text), f"drop-in wrapper for {ref} missing"


def test_abi_version_and_error_strings():
    L = _lib.lib()
    assert L.milab200_abi_version() == 1
    assert b"group_size" in L.milab200_error_string(_lib.E_UNSUPPORTED_GROUP)
    assert L.milab200_error_string(0) == b"success"


def test_argument_errors_are_codes_not_crashes():
    L = _lib.lib()
    nul = None
    assert L.milab200_matvec_decode_bf16_qfp8(nul, nul, nul, nul, nul, 64, 32, nul) == _lib.E_INVALID_ARGUMENT
    one = ctypes.c_void_p(16)        # non-null dummy; argument checks run before any CUDA call
    assert L.milab200_matvec_decode_bf16_qfp4(one, one, one, one, nul, 128, 32, 32, nul) == _lib.E_UNSUPPORTED_GROUP
    assert L.milab200_matvec_decode_bf16_qfp4(one, one, one, one, nul, 192, 32, 128, nul) == _lib.E_BAD_SHAPE
    assert L.milab200_matvec_decode_bf16_qfp8(one, one, one, one, nul, 60, 32, nul) == _lib.E_BAD_SHAPE
    assert L.milab200_quantize_fp4_per_group(one, one, one, 4, 128, 32, one, nul) == _lib.E_UNSUPPORTED_GROUP
    assert L.milab200_quantize_fp4_per_group(one, one, one, 4, 100, 128, one, nul) == _lib.E_BAD_SHAPE
    assert L.milab200_fp4a16_gemm(one, one, one, one, nul, 0, 128, 32, 128, nul) == _lib.E_INVALID_ARGUMENT
    with pytest.raises(_lib.InvalidArgument):
        _lib.check(_lib.E_BAD_SHAPE, "x")
    with pytest.raises(_lib.MilaB200Error):
        _lib.check(_lib.E_UNSUPPORTED_GROUP, "x")      # reference: std::runtime_error (Fp4 .cu:220)


def test_host_mirror_error_conventions_without_gpu():
    import torch
    cfg = LinearConfig(64, 32).withBias(False)
    with pytest.raises(_lib.InvalidArgument):
        LinearConfig(0, 4)
    with pytest.raises(_lib.InvalidArgument):
        Linear("l", cfg, "cpu", PerChannelFp8())        # Linear.ixx:132 device type mismatch
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MilaB200Error):
            Linear("l", cfg, "cuda:0", PerGroupFp4(128))   # no CPU fallback
    assert PerGroupFp4(128).tag == "per_group_fp4_128" and PerChannelFp8().tag == "per_channel_fp8_e4m3"
    assert PerGroupFp4().kIsFp4E2M1 and not PerGroupFp4().kPerChannel and PerChannelFp8().kPerChannel


def test_product_code_never_touches_the_oracle():
    """A product path that routes through the oracle voids parity: nothing under mila_b200/ or
    include/ may mention it."""
    for p in list((ROOT / "mila_b200").rglob("*.py")) + list((ROOT / "mila_b200").rglob("*.cu")) + \
            list((ROOT / "mila_b200").rglob("*.cuh")) + list((ROOT / "include").rglob("*")):
        if p.is_file():
            assert "oracle" not in p.read_text(errors="ignore").lower().replace("oracles'", ""), p
