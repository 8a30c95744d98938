"""Host model of the activation operand of the kind::f8f6f4 kernels (act_split.cuh: split_e4m3x8, used by decode_tc.cu,
prefill_tc.cu and the activation pre-pass): a BF16 activation is written as TWO E4M3 numbers, v = x 2^-e = hi + lo/16.

The claims DESIGN.md §4.1 makes, checked here with the oracle's E4M3 conversion (pinned to the CUDA toolkit's
cvt.rn.satfinite.e4m3x2 by tests/test_oracle.py):
  * with e from the block maximum (absmax 2^-e in [128, 256)), hi = rn_e4m3(v), lo = rn_e4m3(16 (v - hi)) reproduce v
    EXACTLY for every BF16 value with |v| >= 2^-6 (an 8-bit significand is two 4-bit ones);
  * below that the absolute error is < 2^-21 of the block maximum (2^-13 in units of v);
  * the block exponent rule never saturates hi (|v| <= 256 <= 448).
(The device code performs the same steps in packed FP16, every one of them exact; the GPU parity tests check that.)"""
import numpy as np

from oracle import oracle as O


def _e4m3(v: float) -> float:
    L = O.lib()
    return float(L.oracle_e4m3_to_f32(L.oracle_f32_to_e4m3(np.float32(v))))


def _split(v: float):
    hi = _e4m3(v)
    lo = _e4m3(16.0 * (v - hi))
    return hi, lo


def _bf16_values_in(lo_exp: int, hi_exp: int):
    """every positive BF16 value with binary exponent in [lo_exp, hi_exp)"""
    for e in range(lo_exp, hi_exp):
        for m in range(128):
            yield (1.0 + m / 128.0) * 2.0 ** e


def test_block_exponent_rule_keeps_the_maximum_in_128_256():
    """e = exponent(absmax) - 7, clamped to [-100, 100] so that 2^-e, 2^e and the FP32 promotion D 2^e s stay finite:
    block maxima in [2^-93, 2^108) land in [128, 256) and never saturate E4M3.  (Outside — |x| >= 2^108, far beyond
    any activation a BF16 model produces — hi saturates at 448; below 2^-93 only absolute precision is lost.)"""
    for bits in (0x1180, 0x3F80, 0x4049, 0x42C8, 0x7500):          # 2^-92, 1.0, pi, 100.0, 2^107
        e = max(-100, min(100, (bits >> 7) - 127 - 7))
        x = float(O.bf16_bits_to_f32(np.array([bits], np.uint16))[0])
        v = x * 2.0 ** -e
        assert -100 <= e <= 100 and 128.0 <= v < 256.0, (hex(bits), e, v)
    for bits, want in ((0x7F7F, 100), (0x0001, -100), (0x0080, -100)):   # max finite, smallest denormal, min normal
        assert max(-100, min(100, (min(bits, 0x7F7F) >> 7) - 127 - 7)) == want


def test_split_is_exact_down_to_two_to_the_minus_six():
    worst = 0.0
    for v in _bf16_values_in(-6, 8):                                 # 2^-6 <= v < 256: every BF16 significand
        for s in (1.0, -1.0):
            hi, lo = _split(s * v)
            assert hi + lo / 16.0 == s * v, (v, hi, lo)
            worst = max(worst, abs(lo))
    assert worst <= 448.0
    hi, lo = _split(256.0)                                           # the block maximum itself can be 2^8 after rounding up? no: < 256,
    assert hi + lo / 16.0 == 256.0                                   # but the value is representable all the same


def test_error_bound_below_the_exact_range():
    bound = 2.0 ** -13                                               # in units of v; the block maximum is >= 2^7 => < 2^-20 of it
    worst = 0.0
    for v in _bf16_values_in(-20, -6):
        hi, lo = _split(v)
        worst = max(worst, abs(hi + lo / 16.0 - v))
    assert worst <= bound, worst
    assert _split(0.0) == (0.0, 0.0) and _split(-0.0)[0] == 0.0


def test_products_with_weights_are_exact_in_fp32():
    """hi and lo are 4-bit significands, E4M3 / E2M1 weights have <= 4: every product has <= 8 significant bits and a
    128-term sum of them stays far inside FP32's 24 — 'products are exact in the tensor core' needs only this."""
    rng = np.random.default_rng(1)
    L = O.lib()
    w = np.array([L.oracle_e4m3_to_f32(int(b)) for b in rng.integers(0, 256, 128) if (int(b) & 0x7F) != 0x7F][:100], np.float64)
    xs = rng.standard_normal(w.size).astype(np.float32)
    bits = O.f32_to_bf16_bits(xs.reshape(1, -1))
    x = O.bf16_bits_to_f32(bits).reshape(-1).astype(np.float64)
    amax = np.abs(x).max()
    e = int(np.floor(np.log2(amax))) - 7
    hi = np.array([_split(float(v) * 2.0 ** -e)[0] for v in x]); lo = np.array([_split(float(v) * 2.0 ** -e)[1] for v in x])
    d_hi = np.float32(0); d_lo = np.float32(0)
    for k in range(w.size):                                          # FP32 accumulation of exact products
        d_hi = np.float32(d_hi + np.float32(w[k] * hi[k])); d_lo = np.float32(d_lo + np.float32(w[k] * lo[k]))
    exact = float(np.sum(w * x))
    got = (float(d_hi) + float(d_lo) / 16.0) * 2.0 ** e
    assert abs(got - exact) <= 2.0 ** -20 * np.sum(np.abs(w * x)) + 1e-30
