"""CPU tests: pin the oracle (oracle/mila_oracle.c) against every formula / fixture the reference's own
tests hold for this path and against the CUDA toolkit's host E4M3 conversion."""
import ctypes
import math

import numpy as np
import pytest

from oracle import oracle as O
import parity_helpers as H


# ---- E4M3 -----------------------------------------------------------------------------------

def test_e4m3_known_answers():
    f = O.lib().oracle_f32_to_e4m3
    # SURVEY §8c probes of the toolkit conversion
    assert f(17.0) == 0x58 and f(19.0) == 0x5A          # ties to even
    assert f(449.0) == 0x7E and f(1000.0) == 0x7E       # saturate-to-finite
    assert f(-0.0) == 0x80 and f(0.0) == 0x00
    assert f(448.0) == 0x7E and f(-448.0) == 0xFE
    assert f(float("inf")) == 0x7E and f(float("-inf")) == 0xFE
    assert f(float("nan")) == 0x7F
    assert f(2.0 ** -9) == 0x01 and f(2.0 ** -10) == 0x00 and f(1.5 * 2.0 ** -10) == 0x01
    assert f(2.0 ** -6) == 0x08


def test_e4m3_matches_toolkit_on_strided_sweep():
    bad = ctypes.c_uint32(0)
    # every 4099th bit pattern of the full 2^32 space (~1.05 M values) + dense sweeps around the
    # saturation and subnormal boundaries; the exhaustive 2^32 sweep was run once.
    n = O.pin_lib().pin_sweep_e4m3(0, 4099, (1 << 32) // 4099, ctypes.byref(bad))
    assert n == 0, hex(bad.value)
    for centre in (0x43E00000, 0x43E80000, 0x3C800000, 0x3A800000, 0x3B000000, 0x7F800000):
        for sign in (0, 0x80000000):
            n = O.pin_lib().pin_sweep_e4m3((centre | sign) - 4096, 1, 8192, ctypes.byref(bad))
            assert n == 0, hex(bad.value)


def test_e4m3_decode_matches_toolkit_and_reference_test_decoder():
    lib, pin = O.lib(), O.pin_lib()
    for b in range(256):
        a, t = lib.oracle_e4m3_to_f32(b), pin.pin_nv_e4m3_to_f32(b)
        if (b & 0x7F) == 0x7F:
            assert math.isnan(a) and math.isnan(t)
            continue
        assert a == t
        # decodeFp8E4M3 of Linear.Cuda.cpp:929-949
        e, m = (b >> 3) & 0xF, b & 7
        v = m / 8.0 * 2.0 ** -6 if e == 0 else (1 + m / 8.0) * 2.0 ** (e - 7)
        assert a == (-v if b & 0x80 else v)
    # round trip: every finite code converts back to itself
    for b in range(256):
        if (b & 0x7F) != 0x7F:
            assert lib.oracle_f32_to_e4m3(lib.oracle_e4m3_to_f32(b)) == b


# ---- E2M1 -----------------------------------------------------------------------------------

def test_e2m1_ladder_is_half_away_not_rne():
    f = O.lib().oracle_f32_to_e2m1
    # ties go UP in magnitude (CudaFp4WeightQuantization.cu:54-70), unlike cvt.rn e2m1
    ties = {0.25: 1, 0.75: 2, 1.25: 3, 1.75: 4, 2.5: 5, 3.5: 6, 5.0: 7}
    for v, mag in ties.items():
        assert f(v) == mag and f(-v) == (8 | mag)
        below = np.nextafter(np.float32(v), np.float32(0))
        assert f(float(below)) == mag - 1
    assert f(-0.0) == 0 and f(0.0) == 0
    assert f(-1e-30) == 8                      # tiny negative -> "-0"
    assert f(float("nan")) == 7                # all `<` false -> 7, sign test false
    assert f(1e9) == 7 and f(-1e9) == 15
    lut = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
    for nib in range(16):
        v = O.lib().oracle_e2m1_to_f32(nib)
        assert v == (-lut[nib & 7] if nib & 8 else lut[nib & 7])
        if nib != 8:
            assert f(v) == nib                  # exactly representable values round-trip


# ---- quantizers -----------------------------------------------------------------------------

def test_fp8_quantizer_formula_pins():
    w = H.xavier_weights_bf16(48, 256)
    q, s = O.quantize_fp8_per_channel(w)
    wf = O.bf16_bits_to_f32(w)
    absmax = np.max(np.abs(wf), axis=1)
    # TokenEmbedding.Cuda.cpp:569 — scale = row_absmax / 448 (or 1)
    np.testing.assert_array_equal(s, (absmax / np.float32(448.0)).astype(np.float32))
    # Linear.Cuda.cpp:1046 — decode(byte)*scale ~ w within 0.08|w| + 1e-3
    rec = O.dequant_fp8(q, s)
    assert np.all(np.abs(rec - wf) <= 0.08 * np.abs(wf) + 1e-3)
    # TokenEmbedding.Cuda.cpp:555-577 — within 0.07|w| + 0.004*scale
    assert np.all(np.abs(rec - wf) <= 0.07 * np.abs(wf) + 0.004 * s[:, None])
    # the row absmax element always maps to +-448
    idx = np.argmax(np.abs(wf), axis=1)
    assert np.all((q[np.arange(48), idx] & 0x7F) == 0x7E)


def test_fp8_quantizer_corner_cases():
    w = H.adversarial_weights_bf16(16, 256)
    q, s = O.quantize_fp8_per_channel(w)
    assert s[0] == 1.0 and np.all(q[0] == 0)                   # all-zero row
    assert s[11] == 1.0 and np.all(q[11] == 0x80)              # all -0.0 row keeps the sign bit
    assert np.isfinite(s[6]) and q[6, 3] == 0x7F               # NaN ignored by absmax, encodes as NaN
    assert q[6, 4] == 0x7F                                     # negative NaN -> canonical (sign dropped)
    assert np.isinf(s[8]) and q[8, 7] == 0x7F                  # inf*0 = NaN
    assert np.all(q[8, :7] == 0) or np.all((q[8, :7] & 0x7F) == 0)
    assert s[4] == np.float32(np.float32(9984.0) / np.float32(448.0))   # bf16(1e4) = 9984
    assert np.all((q[10] & 0x7F) == 0x7E)


def test_fp4_quantizer_structure_and_packing():
    for g in (64, 128):
        w = H.xavier_weights_bf16(32, 512, seed=7)
        q, s = O.quantize_fp4_per_group(w, g)
        assert q.shape == (32, 256) and q.dtype == np.uint8          # Linear.Cuda.cpp:1055
        assert s.shape == (32, 512 // g) and np.all(np.isfinite(s)) and np.all(s > 0)
        wf = O.bf16_bits_to_f32(w).reshape(32, 512 // g, g)
        np.testing.assert_array_equal(s, (np.max(np.abs(wf), axis=2) / np.float32(6.0)).astype(np.float32))
        # low nibble = even column, high nibble = odd column
        inv = (np.float32(1.0) / s)[:, :, None]
        v = (wf * inv).reshape(32, 512)
        enc = np.vectorize(lambda x: O.lib().oracle_f32_to_e2m1(float(x)), otypes=[np.uint8])(v)
        np.testing.assert_array_equal(q, enc[:, 0::2] | (enc[:, 1::2] << 4))


def test_fp4_exactly_representable_round_trip():
    """The white-box test Mila's Testing.md:84-110 specifies but BACKLOG.md:105-107 lists as unwritten:
    weights that are exactly lut*scale must survive quantize->dequantize unchanged."""
    rng = np.random.default_rng(3)
    lut = np.array([0, .5, 1, 1.5, 2, 3, 4, 6], np.float32)
    nib = rng.integers(0, 16, size=(8, 256))
    nib[:, ::128] = 7                                        # each group contains +6 -> absmax = 6*scale
    vals = lut[nib & 7] * np.where(nib & 8, -1, 1).astype(np.float32)
    scale = np.float32(0.25)                                 # power of two: products exact in BF16
    w = O.f32_to_bf16_bits(vals * scale)
    q, s = O.quantize_fp4_per_group(w, 128)
    assert np.all(s == scale)
    np.testing.assert_array_equal(O.dequant_fp4(q, s, 128), vals * scale)
    dec = np.stack([q & 0xF, q >> 4], axis=-1).reshape(8, 256)
    same = (dec == nib) | ((nib == 8) & (dec == 0))          # -0 is stored as +0 (sign test is `<`)
    assert np.all(same)


def test_fp4_quantizer_corner_cases():
    w = H.adversarial_weights_bf16(16, 256, group=128)
    q, s = O.quantize_fp4_per_group(w, 128)
    assert s[0, 0] == 1.0 and np.all(q[0] == 0)
    assert s[1, 0] == 1.0 and np.all(q[1, :64] == 0)
    # row 2: [6, thresholds..., -thresholds..., -0, 0, -tiny, tiny], scale exactly 1
    assert s[2, 0] == 1.0
    nibs = np.stack([q[2] & 0xF, q[2] >> 4], axis=-1).reshape(-1)[:19]
    assert list(nibs) == [7, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 0, 0, 8, 0]
    # row 3 = row 2 * 0.5 -> scale 0.5, same nibbles
    assert s[3, 0] == 0.5
    nibs3 = np.stack([q[3] & 0xF, q[3] >> 4], axis=-1).reshape(-1)[:19]
    assert list(nibs3) == list(nibs)
    assert s[7, 0] == 1.0 and np.all(q[7, :64] == 0x77)      # all-NaN group: scale 1, nibble 7
    assert np.isinf(s[8, 0])
    with pytest.raises(RuntimeError):
        O.quantize_fp4_per_group(w, 32)
    with pytest.raises(ValueError):
        O.quantize_fp4_per_group(w[:, :192], 128)


def test_reference_fixture_blob_quantizes_consistently():
    # the blob the reference tests feed to loadParameter (Linear.Cuda.cpp:666-677, 793-803)
    w = O.ref_weight_blob(256, 512)
    q8, s8 = O.quantize_fp8_per_channel(w)
    q4, s4 = O.quantize_fp4_per_group(w, 128)
    assert np.all(s8 == s8[0]) and np.all(s4 == s4[0, 0])     # weightValue's absmax is the same in every row/group
    wf = O.bf16_bits_to_f32(w)
    assert s8[0] == np.float32(np.max(np.abs(wf)) / np.float32(448.0))
    assert s4[0, 0] == np.float32(np.max(np.abs(wf)) / np.float32(6.0))


# ---- forward --------------------------------------------------------------------------------

def test_forward_oracle_equals_float64_matmul():
    w = H.xavier_weights_bf16(40, 256)
    x = H.activations_bf16(5, 256)
    bias = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(40)))
    for name in ("fp8", "fp4"):
        if name == "fp8":
            q, s = O.quantize_fp8_per_channel(w); wf = O.dequant_fp8(q, s)
        else:
            q, s = O.quantize_fp4_per_group(w, 128); wf = O.dequant_fp4(q, s, 128)
        y, yf = O.linear_forward_bf16(x, wf, bias)
        ref = O.bf16_bits_to_f32(x).astype(np.float64) @ wf.astype(np.float64).T \
            + O.bf16_bits_to_f32(bias).astype(np.float64)
        np.testing.assert_allclose(yf, ref.astype(np.float32), rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(y, O.f32_to_bf16_bits(yf))


def test_fp4_prefill_vs_decode_fixture_budget():
    """Linear.Cuda.cpp:773-875: M=16 rows of magnitude 1e-8..1e7, K=512, N=256, FP4 g=128.  The oracle's
    batched result is by construction the per-row matvec result; check the fixture is well-formed and
    that BF16 output rounding alone stays far inside the reference budget 1e-1*row_absmax."""
    w = O.ref_weight_blob(256, 512)
    q, s = O.quantize_fp4_per_group(w, 128)
    x = O.f32_to_bf16_bits(O.ref_magnitude_rows(16, 512))
    y, yf = O.linear_forward_fp4(x, q, s, 128)
    for m in range(16):
        y1, yf1 = O.linear_forward_fp4(x[m:m + 1], q, s, 128)
        np.testing.assert_array_equal(y1[0], y[m])
        row_absmax = np.max(np.abs(yf1))
        assert np.all(np.abs(O.bf16_bits_to_f32(y[m]) - yf1[0]) <= 1e-1 * row_absmax)


def test_cpu_linear_forward_paths():
    """Linear.Cpu.cpp:247-360: naive (long double) and unrolled (float) paths vs host ref, tol 1e-4."""
    rng = np.random.default_rng(0)
    for batch in (1, 3, 8, 16):
        X = rng.standard_normal((batch, 64)).astype(np.float32)
        o, i = np.meshgrid(np.arange(32), np.arange(64), indexing="ij")
        W = O.ref_weight_value(o, i); B = O.ref_bias_value(np.arange(32))
        ref = (X.astype(np.float64) @ W.astype(np.float64).T + B).astype(np.float32)
        for path in ("auto", "naive") + (("unrolled",) if batch % 8 == 0 else ()):
            Y = O.cpu_linear_forward(X, W, B, path)
            np.testing.assert_allclose(Y, ref, atol=1e-4, rtol=0)
        np.testing.assert_array_equal(O.cpu_linear_forward(X, W, None, "naive"),
                                      (X.astype(np.longdouble) @ W.astype(np.longdouble).T).astype(np.float32))


def test_w4a8_helper_restatements():
    w = H.xavier_weights_bf16(16, 256)
    q, s = O.quantize_fp4_per_group(w, 128)
    sB = O.compute_fp8_weight_scale(s)
    assert sB == np.float32(np.float32(np.max(s)) * np.float32(np.float32(6.0) / np.float32(448.0)))
    w8 = O.fp4_dequantize_to_fp8(q, s, sB, 128)
    # the group holding the global absmax maps its +-6 nibble to +-448
    assert np.max(w8 & 0x7F) == 0x7E
    w16 = O.fp4_dequantize_to_bf16(q, s, 128)
    np.testing.assert_array_equal(w16, O.f32_to_bf16_bits(O.dequant_fp4(q, s, 128)))
    x = H.activations_bf16(4, 256)
    x8, sA = O.quantize_bf16_to_fp8_per_token(x)
    assert np.all((np.max(x8 & 0x7F, axis=1)) == 0x7E)
    np.testing.assert_array_equal(sA, (np.max(np.abs(O.bf16_bits_to_f32(x)), axis=1) / np.float32(448)).astype(np.float32))


def test_token_embedding_oracle_formula():
    """FP8 tied-table gather: Y = bf16(f32(e4m3) * scale[row]) (TokenEmbedding.Fp8.cu:34-66) against numpy."""
    rng = np.random.default_rng(8)
    V, C = 50, 64
    w = O.f32_to_bf16_bits((rng.standard_normal((V, C)) * 0.1).astype(np.float32))
    q, s = O.quantize_fp8_per_channel(w)
    ids = np.array([0, 49, 7, 7, 13], np.int32)
    y = O.token_embedding_qfp8(ids, q, s)
    deq = O.dequant_fp8(q, s)[ids]
    assert np.array_equal(y, O.f32_to_bf16_bits(deq))


def test_glu_oracle_formulas():
    """GeGLU (tanh) and SwiGLU restatements against float64 formulas, within one BF16 ulp."""
    rng = np.random.default_rng(9)
    x = np.clip(rng.standard_normal((4, 64)) * 1.5, -3.0, 3.0).astype(np.float32)   # 1 + tanh cancels in FP32 beyond that
    xb = O.f32_to_bf16_bits(x); xf = O.bf16_bits_to_f32(xb).astype(np.float64)
    g, u = xf[:, :32], xf[:, 32:]
    ref1 = 0.5 * g * (1 + np.tanh(0.7978845608 * (g + 0.044715 * g ** 3))) * u
    ref2 = g / (1 + np.exp(-g)) * u
    for kind, ref in ((1, ref1), (2, ref2)):
        y = O.bf16_bits_to_f32(O.glu_forward_bf16(xb, kind)).astype(np.float64)
        assert np.all(np.abs(y - ref) <= 2.0 ** -7 * np.abs(ref) + 1e-30)
