"""torch <-> numpy glue for the GPU parity tests (device memory and streams only)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from mila_b200 import _lib
from oracle import oracle as O


def bf16_tensor(bits: np.ndarray, device="cuda:0", pinned=False) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(bits).view(np.int16).copy()).view(torch.bfloat16)
    if device == "cpu":
        return t.pin_memory() if pinned else t
    return t.to(device)


def bits_of(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().cpu().view(torch.int16).numpy().view(np.uint16)


def u8(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().cpu().view(torch.uint8).numpy()


def f32(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().cpu().numpy()


def p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ref_quantize_fp8(w_bits: np.ndarray):
    """The reference's own kernel (oracle/_ref), same inputs."""
    R = O.ref_lib()
    N, K = w_bits.shape
    src = bf16_tensor(w_bits, "cpu", pinned=True)
    st = torch.empty((N, K), dtype=torch.bfloat16, device="cuda")
    q = torch.empty((N, K), dtype=torch.uint8, device="cuda"); s = torch.empty((N,), dtype=torch.float32, device="cuda")
    rc = R.milaref_quantize_fp8_per_channel(p(src), p(q), p(s), ctypes.c_int64(N), ctypes.c_int64(K), p(st),
                                            ctypes.c_void_p(stream()))
    torch.cuda.synchronize(); assert rc == 0
    return u8(q), f32(s)


def ref_quantize_fp4(w_bits: np.ndarray, g: int):
    R = O.ref_lib()
    N, K = w_bits.shape
    src = bf16_tensor(w_bits, "cpu", pinned=True)
    st = torch.empty((N, K), dtype=torch.bfloat16, device="cuda")
    q = torch.empty((N, K // 2), dtype=torch.uint8, device="cuda")
    s = torch.empty((N, K // g), dtype=torch.float32, device="cuda")
    rc = R.milaref_quantize_fp4_per_group(p(src), p(q), p(s), ctypes.c_int64(N), ctypes.c_int64(K), g, p(st),
                                          ctypes.c_void_p(stream()))
    torch.cuda.synchronize(); assert rc == 0
    return u8(q), f32(s)


def ref_matvec(x: torch.Tensor, q: torch.Tensor, s: torch.Tensor, g: int, bias=None) -> torch.Tensor:
    """The reference's own M=1 matvec kernels (g == 0 -> FP8)."""
    R = O.ref_lib()
    K = x.shape[-1]; N = q.shape[0]
    y = torch.empty((N,), dtype=torch.bfloat16, device="cuda")
    if g == 0:
        rc = R.milaref_matvec_decode_bf16_qfp8(p(y), p(x), p(q), p(s), p(bias), K, N, ctypes.c_void_p(stream()))
    else:
        rc = R.milaref_matvec_decode_bf16_qfp4(p(y), p(x), p(q), p(s), p(bias), K, N, g, ctypes.c_void_p(stream()))
    torch.cuda.synchronize(); assert rc == 0
    return y


def generic_gemv(x, q, s, g, bias=None):
    """The one-warp-per-row kernel (route option "decode_generic"): an independent second device implementation."""
    from mila_b200.linear import PerChannelFp8, PerGroupFp4, linear_forward
    _lib.set_option("decode_generic", 1)
    try:
        y = linear_forward(x, q, s, PerChannelFp8() if g == 0 else PerGroupFp4(g), bias)
        torch.cuda.synchronize()
        assert _lib.last_kernel().startswith("gemv_generic_kernel"), _lib.last_kernel()
    finally:
        _lib.set_option("decode_generic", 0)
    return y
