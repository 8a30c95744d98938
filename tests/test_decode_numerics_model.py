"""Host model of the ARITHMETIC of the tcgen05 decode kernels (decode_tc.cu / prefill_tc.cu / decode_mx4.cu), run on
the CPU against the dequantise-then-FP32-GEMM oracle: it shows that the scheme itself — per-(token, 128-k block)
power-of-two exponent, exact operand planes, exact products, FP32 promotion per block with the (row, group) scale, one
BF16 rounding — sits far inside the 1e-2 bar on Xavier inputs AND on the reference's 15-decade row-magnitude fixture
(Linear.Cuda.cpp:822-833), independent of any GPU.  (That the kernels implement this arithmetic is what the GPU
parity tests check.)"""
import numpy as np
import pytest

import parity_helpers as H
from oracle import oracle as O

BLOCK = 128
_E4M3 = None


def _e4m3_round(v: np.ndarray) -> np.ndarray:
    """rn_e4m3 (satfinite) of an array through a 2^16-entry table over the FP16-exact inputs the kernels produce:
    every v here is a BF16 significand times a power of two inside [-256, 256], so a sorted-table lookup is exact."""
    L = O.lib()
    flat = v.reshape(-1)
    out = np.empty_like(flat)
    cache = {}
    for i, x in enumerate(flat):
        x = float(x)
        if x not in cache:
            cache[x] = float(L.oracle_e4m3_to_f32(L.oracle_f32_to_e4m3(np.float32(x))))
        out[i] = cache[x]
    return out.reshape(v.shape)


def _block_exponents(x: np.ndarray) -> np.ndarray:
    """e per (token, block): absmax 2^-e in [128, 256) (act_split.cuh / decode_tc.cu converter)."""
    M, K = x.shape
    am = np.abs(x).reshape(M, K // BLOCK, BLOCK).max(axis=2)
    e = np.where(am > 0, np.floor(np.log2(np.where(am > 0, am, 1.0))).astype(np.int64) - 7, 0)
    return np.clip(e, -100, 100)


def model_f8f6f4(x_bits, wf_unscaled, group_scales, row_scales, bias_bits=None):
    """wf_unscaled [N, K]: the weight codes as numbers (E4M3 values, or E2M1 values for FP4); group_scales [N, K/128]
    (ones for FP8); row_scales [N] (ones for FP4).  Returns BF16 bits [M, N]."""
    x = O.bf16_bits_to_f32(x_bits).astype(np.float64)
    M, K = x.shape
    N = wf_unscaled.shape[0]
    e = _block_exponents(x)
    acc = np.zeros((M, N), np.float32)
    for b in range(K // BLOCK):
        ks = slice(b * BLOCK, (b + 1) * BLOCK)
        v = x[:, ks] * np.exp2(-e[:, b])[:, None]                       # exact: power-of-two scaling
        hi = _e4m3_round(v)
        lo = _e4m3_round(16.0 * (v - hi))
        w = wf_unscaled[:, ks].astype(np.float64)
        d_hi = (hi @ w.T); d_lo = (lo @ w.T)                             # exact products (<= 8 significant bits each); the
        # 128-term sums are rounded to FP32 once here — the tensor core rounds them in its own order, same magnitude
        dv = (d_lo.astype(np.float32) * np.float32(0.0625) + d_hi.astype(np.float32)).astype(np.float32)
        xs = np.exp2(e[:, b]).astype(np.float32)[:, None]
        acc = (acc + (dv * xs) * group_scales[None, :, b].astype(np.float32)).astype(np.float32)
    y = acc * row_scales[None, :].astype(np.float32)
    if bias_bits is not None:
        y = y + O.bf16_bits_to_f32(bias_bits)[None, :]
    return O.f32_to_bf16_bits(y.astype(np.float32))


def model_mxf4_digits(x_bits, wf_unscaled, group_scales):
    """decode_mx4.cu: 16-bit fixed point relative to the block maximum, u = rn(x 2^(15-E)) (the digit planes carry u
    exactly — tests/test_base8_plane_model.py), exact integer sums, FP32 promotion with 2^(E-15) and the group scale."""
    x = O.bf16_bits_to_f32(x_bits).astype(np.float64)
    M, K = x.shape
    N = wf_unscaled.shape[0]
    acc = np.zeros((M, N), np.float32)
    for b in range(K // BLOCK):
        ks = slice(b * BLOCK, (b + 1) * BLOCK)
        am = np.abs(x[:, ks]).max(axis=1)
        E = np.where(am > 0, np.floor(np.log2(np.where(am > 0, am, 1.0))), 0).astype(np.int64)
        E = np.maximum(E, -110)
        u = np.rint(x[:, ks] * np.exp2(15 - E)[:, None])                  # round half to even, like F2I.RN
        d = (u @ wf_unscaled[:, ks].astype(np.float64).T).astype(np.float32)   # |sum| < 2^28: rounded once, like the FP32 accumulator
        acc = (acc + (d * np.exp2(E - 15).astype(np.float32)[:, None]) * group_scales[None, :, b].astype(np.float32)).astype(np.float32)
    return O.f32_to_bf16_bits(acc)


def _fixtures(M, K):
    yield "xavier", H.activations_bf16(M, K, seed=3)
    yield "15-decade rows", O.f32_to_bf16_bits(O.ref_magnitude_rows(M, K))
    sp = O.bf16_bits_to_f32(H.activations_bf16(M, K, seed=4)).copy()
    sp[:, ::7] *= np.float32(2.0 ** -12); sp[:, 5] = 300.0                # outliers and a wide in-block dynamic range
    yield "outliers", O.f32_to_bf16_bits(sp)


@pytest.mark.parametrize("M", [1, 3])
def test_fp8_scheme_meets_the_bar_on_the_cpu(M):
    N, K = 48, 512
    q, s = O.quantize_fp8_per_channel(H.xavier_weights_bf16(N, K, seed=8))
    codes = O.dequant_fp8(q, np.ones(N, np.float32))                       # E4M3 values of the bytes
    bias = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N)))
    for name, xb in _fixtures(M, K):
        want_bits, ref = O.linear_forward_fp8(xb, q, s, bias)
        got = O.bf16_bits_to_f32(model_f8f6f4(xb, codes, np.ones((N, K // BLOCK), np.float32), s, bias))
        err = H.rel_err_rowabs(got, ref)
        assert err <= 2.0 ** -8, (name, err)                               # one BF16 rounding of the result
        ulp_flips = np.mean(O.f32_to_bf16_bits(got.astype(np.float32)) != want_bits)
        assert ulp_flips <= 0.05, (name, ulp_flips)                        # and nearly always the oracle's very bits


@pytest.mark.parametrize("group", [128])
@pytest.mark.parametrize("M", [1, 3])
def test_fp4_schemes_meet_the_bar_on_the_cpu(M, group):
    N, K = 48, 512
    q, s = O.quantize_fp4_per_group(H.xavier_weights_bf16(N, K, seed=9), group)
    codes = O.dequant_fp4(q, np.ones_like(s), group)                       # E2M1 values of the nibbles
    for name, xb in _fixtures(M, K):
        want_bits, ref = O.linear_forward_fp4(xb, q, s, group, None)
        for model in (lambda: model_f8f6f4(xb, codes, s, np.ones(N, np.float32)), lambda: model_mxf4_digits(xb, codes, s)):
            got = O.bf16_bits_to_f32(model())
            err = H.rel_err_rowabs(got, ref)
            assert err <= 2.0 ** -8, (name, err)
            assert np.mean(O.f32_to_bf16_bits(got.astype(np.float32)) != want_bits) <= 0.05, name
