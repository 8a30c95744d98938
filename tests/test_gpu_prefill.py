"""GPU parity: the batched (M > 32) TMA + tcgen05 path (mila_b200/csrc/prefill_tc.cu) through the C-ABI
slots cuda_w8a16_gemm / cuda_fp4a16_gemm.

Bar (BASELINE.json north_star): outputs within max relative error 1e-2 of the dequantise-then-FP32-GEMM
result, measured with the row-abs-normalised metric of SURVEY.md §8d, plus the reference's own BF16 budget
atol 5e-2 + rtol 5e-2 (Linear.Cuda.cpp:121-129).  Small shapes are checked against the CPU oracle; the
BASELINE.json full-size shapes against a torch FP32 GEMM over weights dequantised exactly on the device
(that dequantisation is itself checked against the oracle first), plus size-independent properties."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import (PerChannelFp8, PerGroupFp4, linear_forward, quantize_fp4_per_group,
                              quantize_fp8_per_channel)
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2
POLICIES = [PerChannelFp8(), PerGroupFp4(128)]
IDS = ["fp8", "fp4g128"]
E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0]


def _quantize(policy, w_dev):
    if isinstance(policy, PerChannelFp8):
        return quantize_fp8_per_channel(w_dev)
    return quantize_fp4_per_group(w_dev, policy.kQuantizationGroupSize)


def _dequant_device(policy, q, s):
    """Exact FP32 dequantisation with torch ops (Policies.ixx:89-95, Bf16.cu:20-25)."""
    if isinstance(policy, PerChannelFp8):
        return q.view(torch.float8_e4m3fn).float() * s[:, None]
    g = policy.kQuantizationGroupSize
    lut = torch.tensor(E2M1, dtype=torch.float32, device=q.device)
    lo = lut[(q & 0xF).long()]; hi = lut[(q >> 4).long()]
    w = torch.stack((lo, hi), dim=-1).reshape(q.shape[0], -1)
    return w * s.repeat_interleave(g, dim=1)


def _torch_ref(policy, x, q, s, bias=None):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = x.float() @ _dequant_device(policy, q, s).t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    if bias is not None:
        y = y + bias.float()
    return y


def _check(y, ref, tol=TOL):
    y = np.asarray(y, np.float64); ref = np.asarray(ref, np.float64)
    err = H.rel_err_rowabs(y, ref)
    assert err <= tol, f"row-abs relative error {err:.4g} > {tol}"
    assert np.all(np.abs(y - ref) <= 5e-2 + 5e-2 * np.abs(ref))


def _is_prefill_kernel():
    return _lib.last_kernel().startswith("prefill_tc_kernel")


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
@pytest.mark.parametrize("N,K,M,bias", [(128, 128, 64, False), (256, 512, 33, True), (200, 384, 130, True),
                                        (384, 1024, 257, False), (130, 640, 128, False)])
def test_prefill_matches_oracle(policy, N, K, M, bias):
    w = H.xavier_weights_bf16(N, K, seed=1234 + M)
    x = H.activations_bf16(M, K, seed=99 + M)
    b = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N))) if bias else None
    q, s = _quantize(policy, G.bf16_tensor(w, "cuda"))
    if isinstance(policy, PerChannelFp8):
        _, yf = O.linear_forward_fp8(x, G.u8(q), G.f32(s), b)
    else:
        _, yf = O.linear_forward_fp4(x, G.u8(q), G.f32(s), 128, b)
    xd = G.bf16_tensor(x, "cuda"); bd = None if b is None else G.bf16_tensor(b, "cuda")
    y = linear_forward(xd, q, s, policy, bd)
    torch.cuda.synchronize()
    assert _is_prefill_kernel(), _lib.last_kernel()
    _check(y.float().cpu().numpy(), yf)
    # the device dequantisation used by the full-size tests is the oracle's, bit for bit
    wf = _dequant_device(policy, q, s).cpu().numpy()
    wo = O.dequant_fp8(G.u8(q), G.f32(s)) if isinstance(policy, PerChannelFp8) else O.dequant_fp4(G.u8(q), G.f32(s), 128)
    np.testing.assert_array_equal(wf, wo)
    # and the torch FP32 reference agrees with the oracle GEMM far inside the gate
    yt = _torch_ref(policy, xd, q, s, bd).cpu().numpy()
    assert H.rel_err_rowabs(yt, yf) <= 1e-4


@pytest.mark.parametrize("name,policy,N,K,M", [
    ("llama8b_gate_fp8", PerChannelFp8(), 14336, 4096, 2048),
    ("llama8b_down_fp8", PerChannelFp8(), 4096, 14336, 2048),
    ("gemma_gate_up_fp4", PerGroupFp4(128), 30720, 3840, 2048),
    ("gemma_down_fp4", PerGroupFp4(128), 3840, 15360, 2048),
    ("gemma_qkv_fp4", PerGroupFp4(128), 8192, 3840, 512),
    ("gemma_o_fp4_ragged_m", PerGroupFp4(128), 3840, 4096, 333),
    ("gemma_o_fp8_ragged_m", PerChannelFp8(), 3840, 4096, 1000),
    # mid-size M on layers with few row tiles: split-K (P = 4) with the deterministic ticket fix-up
    ("llama8b_down_fp8_m128_splitk", PerChannelFp8(), 4096, 14336, 128),
    ("gemma_down_fp4_m64_splitk", PerGroupFp4(128), 3840, 15360, 64),
    ("gemma_o_fp4_m256_splitk", PerGroupFp4(128), 3840, 4096, 256),
    ("gemma_o_fp8_m200_splitk", PerChannelFp8(), 3840, 4096, 200),
])
def test_full_size_config_shapes(name, policy, N, K, M):
    """BASELINE.json configs[3]: 2048-token batch at the Gemma 4 12B / Llama-3.1-8B layer shapes."""
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    w = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    x = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    q, s = _quantize(policy, w)
    y = linear_forward(x, q, s, policy)
    torch.cuda.synchronize()
    assert _is_prefill_kernel(), _lib.last_kernel()
    ref = _torch_ref(policy, x, q, s)
    yf = y.float()
    row_abs = ref.abs().amax(dim=1, keepdim=True)
    den = torch.maximum(ref.abs(), 1e-2 * row_abs)
    err = float(((yf - ref).abs() / den).max())
    assert err <= TOL, f"{name}: row-abs relative error {err:.4g}"
    assert bool(((yf - ref).abs() <= 5e-2 + 5e-2 * ref.abs()).all())
    # deterministic: a second forward gives identical bits (the reference compares forwards with EXPECT_EQ)
    y2 = linear_forward(x, q, s, policy)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_prefill_agrees_with_decode_rows(policy):
    """Linear.Cuda.cpp:773 relation (batched vs per-row decode), which the reference holds to
    1e-1*row_absmax; both our paths use exact weights and FP32 accumulation, so we hold 2^-7 (one BF16
    ulp of the row maximum)."""
    N, K, M = 512, 1024, 48
    wb = O.ref_weight_blob(N, K)
    q, s = _quantize(policy, G.bf16_tensor(wb, "cuda"))
    rows = O.ref_magnitude_rows(16, K)                                   # 1e-8 .. 1e7
    xb = O.f32_to_bf16_bits(np.concatenate([rows, rows[::-1], rows * np.float32(0.37)], axis=0))
    xd = G.bf16_tensor(xb, "cuda")
    yb = linear_forward(xd, q, s, policy)
    torch.cuda.synchronize()
    assert _is_prefill_kernel()
    for m in range(M):
        y1 = linear_forward(xd[m:m + 1], q, s, policy)
        a = y1.float().cpu().numpy().reshape(-1); b = yb[m].float().cpu().numpy()
        assert np.all(np.abs(a - b) <= 2.0 ** -7 * np.max(np.abs(a)) + 1e-30), m


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_special_activation_values(policy):
    K, N, M = 512, 256, 40
    xf = O.bf16_bits_to_f32(H.activations_bf16(M, K)).copy()
    xf[0, :] = 0.0                       # all-zero token
    xf[1, :] *= 1e-30                    # tiny token
    xf[2, :] *= 3e30                     # huge token
    xf[3, 5] = 1e4                       # one outlier: the rest of the row sits 2^-13 below the token maximum
    xf[4, :] = 0.0; xf[4, 77] = -3.0     # a single non-zero
    xb = O.f32_to_bf16_bits(xf)
    w = H.xavier_weights_bf16(N, K)
    q, s = _quantize(policy, G.bf16_tensor(w, "cuda"))
    if isinstance(policy, PerChannelFp8):
        _, yf = O.linear_forward_fp8(xb, G.u8(q), G.f32(s), None)
    else:
        _, yf = O.linear_forward_fp4(xb, G.u8(q), G.f32(s), 128, None)
    y = linear_forward(G.bf16_tensor(xb, "cuda"), q, s, policy)
    torch.cuda.synchronize()
    assert _is_prefill_kernel()
    _check(y.float().cpu().numpy(), yf)
    assert bool((y[0] == 0).all())


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_nonfinite_activation_poisons_only_its_token(policy):
    K, N, M = 256, 128, 34
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    x[7, 3] = float("nan"); x[9, 200] = float("inf")
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = _quantize(policy, w)
    y = linear_forward(x, q, s, policy).float()
    torch.cuda.synchronize()
    assert _is_prefill_kernel()
    bad = ~torch.isfinite(y).all(dim=1)
    assert bool(bad[7]) and bool(bad[9]) and int(bad.sum()) == 2


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_linearity_and_column_pick(policy):
    """f(4x) == 4 f(x) exactly (power-of-two scaling is exact in every stage), and x = e_k rows pick
    out columns of the dequantised weights exactly."""
    N, K, M = 1024, 2048, 192
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = _quantize(policy, w)
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    y1 = linear_forward(x, q, s, policy).clone()
    y2 = linear_forward(x * 4.0, q, s, policy)
    torch.cuda.synchronize()
    assert _is_prefill_kernel()
    assert torch.equal(y1.float() * 4.0, y2.float())
    e = torch.zeros((M, K), device="cuda", dtype=torch.bfloat16)
    cols = torch.arange(M, device="cuda") * 7 % K
    e[torch.arange(M, device="cuda"), cols] = 1.0
    ye = linear_forward(e, q, s, policy)
    wf = _dequant_device(policy, q, s)
    assert torch.equal(ye, wf[:, cols].t().contiguous().to(torch.bfloat16))


def test_graph_capture_after_reserve():
    N, K, M = 512, 512, 300
    L = _lib.lib()
    _lib.check(L.milab200_reserve_prefill(M, K), "reserve")
    pol = PerChannelFp8()
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = _quantize(pol, w)
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    eager = linear_forward(x, q, s, pol).clone()
    out = torch.empty_like(eager)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            linear_forward(x, q, s, pol, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_token_blocked_fallback_still_matches(policy):
    """With the tcgen05 prefill kernel switched off the same call takes the token-blocked decode
    kernels; both agree to BF16 rounding."""
    N, K, M = 384, 512, 70
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = _quantize(policy, w)
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    y1 = linear_forward(x, q, s, policy)
    torch.cuda.synchronize()
    assert _is_prefill_kernel()
    _lib.set_option("prefill_tc", 0)
    try:
        y2 = linear_forward(x, q, s, policy)
        torch.cuda.synchronize()
        assert not _is_prefill_kernel()
    finally:
        _lib.set_option("prefill_tc", 1)
    assert H.rel_err_rowabs(y1.float().cpu().numpy(), y2.float().cpu().numpy()) <= 1e-2


@pytest.mark.parametrize("M,N,K,bias", [(256, 512, 1024, False), (300, 1000, 2048, True), (2048, 14336, 4096, False)])
def test_one_plane_fp8_rate_mode(M, N, K, bias):
    """Opt-in FP8-rate mode (route option "prefill_act_planes" = 1): ONE per-token-scaled E4M3 activation plane — the
    reference's own W4A8 activation format (CudaFp8Prefill.cu:116-165) — against raw E4M3 weights.  Checked (a) against a
    bit-faithful emulation of that pipeline (the library's bit-exact per-token quantizer, FP32 GEMM over the exact FP8
    values, y = acc * sA[m] * scale[n] + bias): only FP32 summation order and one BF16 rounding may differ; (b) against the
    exact FP32 reference at the reference's own gate for this format, 1e-1 of the row maximum (Linear.Cuda.cpp:749-773)."""
    policy = PerChannelFp8()
    w = G.bf16_tensor(H.xavier_weights_bf16(N, K, seed=N % 97), "cuda")
    q, s = quantize_fp8_per_channel(w)
    x = G.bf16_tensor(H.activations_bf16(M, K, seed=M % 89), "cuda")
    b = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16) if bias else None
    L = _lib.lib()
    _lib.set_option("prefill_act_planes", 1)
    try:
        y = linear_forward(x, q, s, policy, b)
        torch.cuda.synchronize()
        assert _lib.last_kernel().endswith(",a8>"), _lib.last_kernel()
        y2 = linear_forward(x, q, s, policy, b); torch.cuda.synchronize()
        assert torch.equal(y, y2)
    finally:
        _lib.set_option("prefill_act_planes", 2)
    # (a) emulation of the quantized pipeline
    x8 = torch.empty((M, K), dtype=torch.uint8, device="cuda"); sA = torch.empty((M,), dtype=torch.float32, device="cuda")
    _lib.check(L.milab200_quantize_bf16_to_fp8_per_token(G.p(x8), G.p(sA), G.p(x), M, K, ctypes.c_void_p(G.stream())), "q")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        acc = x8.view(torch.float8_e4m3fn).float() @ q.view(torch.float8_e4m3fn).float().t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    emu = acc * sA[:, None] * s[None, :] + (b.float() if b is not None else 0.0)
    assert H.rel_err_rowabs(y.float().cpu().numpy(), emu.cpu().numpy()) <= 1e-2
    # (b) the exact result, at the reference's gate for one-plane activations
    ref = _torch_ref(policy, x, q, s, b)
    row_abs = ref.abs().amax(dim=1, keepdim=True)
    assert bool(((y.float() - ref).abs() <= 1e-1 * row_abs).all())
    # default mode is untouched: exact split, 1e-2
    y_exact = linear_forward(x, q, s, policy, b); torch.cuda.synchronize()
    assert not _lib.last_kernel().endswith(",a8>")
    _check(y_exact.float().cpu().numpy(), ref.cpu().numpy())


# ---- FP4 weights, summed planes (prefill_tc.cu SP): hi and lo = rn(v - hi) accumulate into the same TMEM columns ------------
@pytest.mark.parametrize("N,K,M,bias", [(256, 512, 256, False), (512, 1024, 300, True), (300, 640, 513, True), (1024, 128, 256, False)])
def test_fp4_summed_planes_matches_oracle(N, K, M, bias):
    """Forced on small shapes (option 2) so the CPU oracle can check it: 256-token tiles incl. ragged M and ragged N, bias,
    one-group K; then the two FP4 modes against each other (both exact-weight, FP32-accumulate: one BF16 ulp of the row
    maximum) and run-to-run determinism."""
    policy = PerGroupFp4(128)
    w = H.xavier_weights_bf16(N, K, seed=77 + N)
    x = H.activations_bf16(M, K, seed=5 + M)
    b = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N))) if bias else None
    q, s = _quantize(policy, G.bf16_tensor(w, "cuda"))
    _, yf = O.linear_forward_fp4(x, G.u8(q), G.f32(s), 128, b)
    xd = G.bf16_tensor(x, "cuda"); bd = None if b is None else G.bf16_tensor(b, "cuda")
    _lib.set_option("prefill_fp4_sum", 2)
    try:
        y = linear_forward(xd, q, s, policy, bd).clone()
        torch.cuda.synchronize()
        assert _lib.last_kernel().endswith(",sum>"), _lib.last_kernel()
        y2 = linear_forward(xd, q, s, policy, bd); torch.cuda.synchronize()
        assert torch.equal(y, y2)
        _lib.set_option("prefill_fp4_sum", 0)
        y0 = linear_forward(xd, q, s, policy, bd).clone(); torch.cuda.synchronize()
        assert _lib.last_kernel().startswith("prefill_tc_kernel") and not _lib.last_kernel().endswith(",sum>")
    finally:
        _lib.set_option("prefill_fp4_sum", 1)
    _check(y.float().cpu().numpy(), yf)
    a = y.float().cpu().numpy(); c = y0.float().cpu().numpy()
    assert np.all(np.abs(a - c) <= 2.0 ** -7 * np.max(np.abs(c), axis=1, keepdims=True) + 1e-30)


def test_fp4_summed_planes_wide_dynamic_range():
    """Tokens spanning 15 decades and one-outlier tokens (everything else 2^-13 below the token maximum): the summed-plane
    split is exact down to 2^-10 of the token maximum and loses < 2^-17 of it below — far inside the gate."""
    policy = PerGroupFp4(128)
    N, K, M = 512, 1024, 256
    rows = O.ref_magnitude_rows(16, K)                                   # 1e-8 .. 1e7
    xf = np.concatenate([rows] * 16, axis=0).astype(np.float32)
    xf[3, :] = O.bf16_bits_to_f32(H.activations_bf16(1, K, seed=3))[0]; xf[3, 5] = 1e4
    xf[4, :] = 0.0
    xb = O.f32_to_bf16_bits(xf)
    q, s = _quantize(policy, G.bf16_tensor(O.ref_weight_blob(N, K), "cuda"))
    _, yf = O.linear_forward_fp4(xb, G.u8(q), G.f32(s), 128, None)
    _lib.set_option("prefill_fp4_sum", 2)
    try:
        y = linear_forward(G.bf16_tensor(xb, "cuda"), q, s, policy).clone()
        torch.cuda.synchronize()
        assert _lib.last_kernel().endswith(",sum>"), _lib.last_kernel()
    finally:
        _lib.set_option("prefill_fp4_sum", 1)
    _check(y.float().cpu().numpy(), yf)
    assert bool((y[4] == 0).all())


@pytest.mark.parametrize("M,N,K,bias", [(256, 512, 1024, False), (300, 1000, 2048, True), (2048, 15360, 3840, False)])
def test_one_plane_fp4_w4a8_mode(M, N, K, bias):
    """Opt-in W4A8 mode for FP4 weights: ONE per-token-scaled E4M3 activation plane — the reference's own batched FP4 path
    quantizes activations exactly like this (cuda_quantize_bf16_to_fp8_per_token, LIN/CudaLinearOp.ixx:660-714) — against the
    raw E2M1 weights with the per-group FP32 promotion.  Checked (a) against a bit-faithful emulation (the library's bit-exact
    per-token quantizer, FP32 GEMM over exact FP8 x dequantised FP4, y = acc * sA[m] + bias) and (b) against the exact FP32
    reference at the reference's gate for this format, 1e-1 of the row maximum (Linear.Cuda.cpp:749-773)."""
    policy = PerGroupFp4(128)
    w = G.bf16_tensor(H.xavier_weights_bf16(N, K, seed=N % 97), "cuda")
    q, s = quantize_fp4_per_group(w, 128)
    x = G.bf16_tensor(H.activations_bf16(M, K, seed=M % 89), "cuda")
    b = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16) if bias else None
    L = _lib.lib()
    _lib.set_option("prefill_act_planes", 1)
    try:
        y = linear_forward(x, q, s, policy, b).clone()
        torch.cuda.synchronize()
        assert _lib.last_kernel().endswith(",a8>"), _lib.last_kernel()
        y2 = linear_forward(x, q, s, policy, b); torch.cuda.synchronize()
        assert torch.equal(y, y2)
    finally:
        _lib.set_option("prefill_act_planes", 2)
    x8 = torch.empty((M, K), dtype=torch.uint8, device="cuda"); sA = torch.empty((M,), dtype=torch.float32, device="cuda")
    _lib.check(L.milab200_quantize_bf16_to_fp8_per_token(G.p(x8), G.p(sA), G.p(x), M, K, ctypes.c_void_p(G.stream())), "q")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        acc = x8.view(torch.float8_e4m3fn).float() @ _dequant_device(policy, q, s).t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    emu = acc * sA[:, None] + (b.float() if b is not None else 0.0)
    assert H.rel_err_rowabs(y.float().cpu().numpy(), emu.cpu().numpy()) <= 1e-2
    ref = _torch_ref(policy, x, q, s, b)
    row_abs = ref.abs().amax(dim=1, keepdim=True)
    assert bool(((y.float() - ref).abs() <= 1e-1 * row_abs).all())
    # default mode is untouched: exact planes, 1e-2
    y_exact = linear_forward(x, q, s, policy, b); torch.cuda.synchronize()
    assert not _lib.last_kernel().endswith(",a8>")
    _check(y_exact.float().cpu().numpy(), ref.cpu().numpy())
