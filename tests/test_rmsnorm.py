"""RMSNorm -> Linear (second half).  GPU: milab200_rmsnorm_forward_bf16 equals the reference's
cuda_rmsnorm_forward_bf16 (compiled unmodified into oracle/_ref) bit for bit; the fused milab200_rmsnorm_*_gemm equals the
two-kernel sequence (reference RMSNorm kernel -> our Linear) bit for bit on every route (decode converters, 9..16-token
pre-pass, batched split kernel, stand-alone fallback), Gemma's (1 + weight) offset included.  CPU: the C restatement
against a float64 evaluation (one BF16 ulp: the device uses MUFU.RSQ) and, when present, the golden vectors of the
reference kernel."""
import ctypes
from pathlib import Path

import numpy as np
import pytest

import parity_helpers as H
import parity_helpers as H_
from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def _ulp_close_bf16(a_bits, b_bits, ulps=1):
    a = a_bits.astype(np.int32); b = b_bits.astype(np.int32)
    # map sign-magnitude BF16 bit patterns onto a monotone integer line
    a = np.where(a & 0x8000, 0x8000 - a, a); b = np.where(b & 0x8000, 0x8000 - b, b)
    return int(np.abs(a - b).max()) <= ulps


@pytest.mark.parametrize("offset", [0.0, 1.0])
@pytest.mark.parametrize("with_bias", [False, True])
def test_oracle_rmsnorm_matches_float64(offset, with_bias):
    M, K = 5, 768
    x = H.activations_bf16(M, K, seed=3)
    w = O.f32_to_bf16_bits((np.random.default_rng(1).standard_normal(K) * 0.2).astype(np.float32))
    b = O.f32_to_bf16_bits((np.random.default_rng(2).standard_normal(K) * 0.1).astype(np.float32)) if with_bias else None
    y, rstd = O.rmsnorm_forward_bf16(x, w, b, 1e-6, offset)
    xf = O.bf16_bits_to_f32(x).astype(np.float64)
    r = 1.0 / np.sqrt((xf * xf).mean(axis=1, keepdims=True) + 1e-6)
    ref = xf * r * (O.bf16_bits_to_f32(w).astype(np.float64) + offset) + (O.bf16_bits_to_f32(b).astype(np.float64) if with_bias else 0.0)
    assert np.allclose(rstd, r[:, 0], rtol=1e-6)
    assert _ulp_close_bf16(y, O.f32_to_bf16_bits(ref.astype(np.float32)), 1)


def test_oracle_rmsnorm_against_reference_kernel_golden():
    f = GOLD / "rmsnorm_ref.npz"
    if not f.exists():
        pytest.skip("golden vectors of the reference RMSNorm kernel not generated yet (tools/make_golden.py on a GPU box)")
    g = np.load(f)
    for tag in ("plain", "gemma"):
        y, _ = O.rmsnorm_forward_bf16(g[f"{tag}_x"], g[f"{tag}_w"], g[f"{tag}_b"] if f"{tag}_b" in g else None,
                                      float(g[f"{tag}_eps"]), float(g[f"{tag}_off"]))
        assert _ulp_close_bf16(y, g[f"{tag}_y"], 1), tag


gpu = pytest.mark.gpu


def _ref_rmsnorm(xd, wd, bd, eps, off):
    import torch
    import gpu_util as G
    R = O.ref_lib()
    M, K = xd.shape
    y = torch.empty_like(xd)
    rc = R.milaref_rmsnorm_forward_bf16(G.p(y), None, G.p(xd), G.p(wd), G.p(bd), M, 1, K, ctypes.c_float(eps), ctypes.c_float(off),
                                        ctypes.c_void_p(G.stream()))
    torch.cuda.synchronize()
    assert rc == 0
    return y


@gpu
@pytest.mark.skipif(not O.ref_lib_path().exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("M,K", [(1, 4096), (7, 3840), (16, 512), (40, 1024), (3, 200)])
@pytest.mark.parametrize("off,with_w,with_b", [(0.0, True, False), (1.0, True, True), (0.0, False, False)])
def test_standalone_rmsnorm_equals_reference_kernel_bit_for_bit(M, K, off, with_w, with_b):
    import torch
    import gpu_util as G
    from mila_b200.linear import rmsnorm_forward
    xd = G.bf16_tensor(H.activations_bf16(M, K, seed=M + K), "cuda")
    wd = (torch.randn(K, device="cuda") * 0.3).to(torch.bfloat16) if with_w else None
    bd = (torch.randn(K, device="cuda") * 0.1).to(torch.bfloat16) if with_b else None
    got = rmsnorm_forward(xd, wd, bd, 1e-6, off)
    torch.cuda.synchronize()
    assert torch.equal(got, _ref_rmsnorm(xd, wd, bd, 1e-6, off))


@gpu
@pytest.mark.skipif(not O.ref_lib_path().exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("policy_name", ["fp8", "fp4g128", "fp4g64"])
@pytest.mark.parametrize("M,N,K,route", [(1, 3840, 4096, "decode_tc_kernel"), (5, 256, 512, "decode_tc_kernel"),
                                         (8, 14336, 4096, "decode_tc_kernel"), (12, 3840, 4096, "decode_tc_kernel"),
                                         (16, 200, 1152, "decode_tc_kernel"), (48, 512, 1024, "prefill_tc_kernel"),
                                         (300, 1000, 2048, "prefill_tc_kernel"), (3, 64, 192, "gemv")])
def test_fused_rmsnorm_linear_equals_two_kernel_sequence_bit_for_bit(policy_name, M, N, K, route):
    import torch
    import gpu_util as G
    from mila_b200 import _lib
    from mila_b200.linear import (PerChannelFp8, PerGroupFp4, linear_forward, quantize_fp4_per_group, quantize_fp8_per_channel,
                                  rmsnorm_linear_forward)
    policy = {"fp8": PerChannelFp8(), "fp4g128": PerGroupFp4(128), "fp4g64": PerGroupFp4(64)}[policy_name]
    if policy_name != "fp8" and K % policy.kQuantizationGroupSize != 0:
        pytest.skip("K not a multiple of the group size")
    w = G.bf16_tensor(H.xavier_weights_bf16(N, K, seed=N), "cuda")
    q, s = quantize_fp8_per_channel(w) if policy_name == "fp8" else quantize_fp4_per_group(w, policy.kQuantizationGroupSize)
    xd = G.bf16_tensor(H.activations_bf16(M, K, seed=M), "cuda") * 3.0
    gamma = (torch.randn(K, device="cuda") * 0.2).to(torch.bfloat16)
    beta = (torch.randn(K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    for (gw, gb, off) in ((gamma, None, 1.0), (gamma, beta, 0.0)):           # Gemma (1 + w) without bias; plain with bias
        fused = rmsnorm_linear_forward(xd, gw, gb, 1e-6, off, q, s, policy, bias)
        torch.cuda.synchronize()
        k_fused = _lib.last_kernel()
        normed = _ref_rmsnorm(xd, gw, gb, 1e-6, off)
        # FP4 g = 128 at M <= 2: both the fused call and the plain Linear take the packed-nibble kernel (decode_mx4.cu)
        seq = linear_forward(normed, q, s, policy, bias)
        torch.cuda.synchronize()
        want = "decode_mx4_kernel" if (policy_name == "fp4g128" and M <= 2 and route == "decode_tc_kernel") else route
        if policy_name != "fp4g64" and route != "gemv":
            assert k_fused.startswith(want), k_fused
            assert _lib.last_kernel().startswith(want), _lib.last_kernel()
        assert torch.equal(fused, seq), (k_fused, _lib.last_kernel())


@gpu
@pytest.mark.skipif(not O.ref_lib_path().exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("policy_name", ["fp8", "fp4g128"])
@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("M,H,K", [(1, 15360, 3840), (2, 14336, 4096), (4, 14336, 4096), (8, 2048, 1024), (13, 15360, 3840),
                                   (16, 14336, 4096), (3, 100, 192), (40, 512, 1024), (512, 2048, 1024), (2048, 4096, 512)])
def test_fused_rmsnorm_gate_up_glu_equals_three_kernel_sequence_bit_for_bit(policy_name, kind, M, H, K):
    """milab200_rmsnorm_*_gemm_glu == reference RMSNorm kernel (oracle/_ref) -> gate|up Linear -> our bit-exact GeGLU / SwiGLU
    kernel, on the one-launch decode routes (packed-nibble FP4 at M <= 2, tcgen05 plane kernels, 9..16-token pre-pass) and
    on the fallback sequence (odd shapes, M > 16); Gemma's (1 + w) offset for GeGLU, plain weights for SwiGLU."""
    import torch
    import gpu_util as G
    from mila_b200 import _lib
    from mila_b200.linear import (PerChannelFp8, PerGroupFp4, glu_forward, linear_forward, quantize_fp4_per_group,
                                  quantize_fp8_per_channel, rmsnorm_linear_glu_forward)
    policy = {"fp8": PerChannelFp8(), "fp4g128": PerGroupFp4(128)}[policy_name]
    if policy_name != "fp8" and K % 128 != 0:
        pytest.skip("K not a multiple of the group size")
    w = G.bf16_tensor(H_.xavier_weights_bf16(2 * H, K, seed=H % 101), "cuda")
    q, s = quantize_fp8_per_channel(w) if policy_name == "fp8" else quantize_fp4_per_group(w, 128)
    xd = G.bf16_tensor(H_.activations_bf16(M, K, seed=M), "cuda") * 2.0
    gamma = (torch.randn(K, device="cuda") * 0.2).to(torch.bfloat16)
    off = 1.0 if kind == 1 else 0.0
    if kind == 2: gamma = (gamma.float() + 1.0).to(torch.bfloat16)
    before = _lib.launch_count()
    fused = rmsnorm_linear_glu_forward(xd, gamma, None, 1e-6, off, q, s, policy, kind)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - before
    k_fused = _lib.last_kernel()
    normed = _ref_rmsnorm(xd, gamma, None, 1e-6, off)
    gu = linear_forward(normed, q, s, policy)
    seq = glu_forward(gu, kind)
    torch.cuda.synchronize()
    assert torch.equal(fused, seq), k_fused
    if M >= 512:
        # batched: the norm in the activation pre-pass, the activation in the GEMM epilogue — two launches, no intermediates
        if policy_name == "fp8":
            assert launches == 2 and k_fused.endswith("glu>"), (launches, k_fused)
    if M <= 8 and K % 128 == 0 and H >= 14336:
        assert launches == 1, (launches, k_fused)                 # ONE launch: norm prologue + Linear + GLU epilogue
        assert k_fused.startswith("decode_mx4_kernel" if (policy_name == "fp4g128" and M <= 2) else "decode_tc_kernel"), k_fused


@gpu
def test_rmsnorm_linear_argument_errors():
    from mila_b200 import _lib
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.milab200_rmsnorm_w8a16_gemm(None, None, one, None, None, 1e-6, 0.0, one, one, None, 1, 128, 128, None) == _lib.E_INVALID_ARGUMENT
    assert L.milab200_rmsnorm_fp4a16_gemm(one, None, one, None, None, 1e-6, 0.0, one, one, None, 1, 128, 128, 32, None) == _lib.E_UNSUPPORTED_GROUP
    assert L.milab200_rmsnorm_forward_bf16(None, None, one, None, None, 1, 1, 8, 1e-6, 0.0, None) == _lib.E_INVALID_ARGUMENT


@gpu
@pytest.mark.parametrize("policy_name", ["fp8", "fp4g128"])
@pytest.mark.parametrize("M", [1, 2, 8, 16])
def test_fast_reduction_option_stays_within_a_bf16_ulp_of_the_reference_order(policy_name, M):
    """Opt-in tree-order sum of squares (milab200_set_option rmsnorm_fast_reduction): rstd differs from the reference's
    lane-strided FMA chains in the last FP32 bits only, so a normalised activation can move by at most one BF16 ulp and the
    fused MLP front half stays far inside the gate; the default (reference order, bit-identical) is untouched."""
    import torch
    import gpu_util as G
    from mila_b200 import _lib
    from mila_b200.linear import PerChannelFp8, PerGroupFp4, quantize_fp4_per_group, quantize_fp8_per_channel, rmsnorm_linear_glu_forward
    policy = {"fp8": PerChannelFp8(), "fp4g128": PerGroupFp4(128)}[policy_name]
    H, K = 15360, 3840
    w = G.bf16_tensor(H_.xavier_weights_bf16(2 * H, K, seed=11), "cuda")
    q, s = quantize_fp8_per_channel(w) if policy_name == "fp8" else quantize_fp4_per_group(w, 128)
    xd = G.bf16_tensor(H_.activations_bf16(M, K, seed=M), "cuda") * 2.0
    gamma = (torch.randn(K, device="cuda") * 0.2).to(torch.bfloat16)
    exact = rmsnorm_linear_glu_forward(xd, gamma, None, 1e-6, 1.0, q, s, policy, 1).clone()
    _lib.set_option("rmsnorm_fast_reduction", 1)
    try:
        fast = rmsnorm_linear_glu_forward(xd, gamma, None, 1e-6, 1.0, q, s, policy, 1).clone()
        fast2 = rmsnorm_linear_glu_forward(xd, gamma, None, 1e-6, 1.0, q, s, policy, 1).clone()
    finally:
        _lib.set_option("rmsnorm_fast_reduction", 0)
    again = rmsnorm_linear_glu_forward(xd, gamma, None, 1e-6, 1.0, q, s, policy, 1)
    torch.cuda.synchronize()
    assert torch.equal(exact, again) and torch.equal(fast, fast2)
    a, b = exact.float(), fast.float()
    row_abs = a.abs().amax(dim=1, keepdim=True)
    assert float(((a - b).abs() / row_abs).max()) <= 2.0 ** -7            # one BF16 ulp of the row maximum
