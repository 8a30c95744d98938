"""GPU parity: decode GEMV (M = 1..16) and the batched slots through the C-ABI vs the dequantise-
then-FP32-GEMM oracle.  Gate: max |y - ref| / max(|ref|, 1e-2*row_absmax) <= 1e-2,
plus the reference's own BF16 budget atol 5e-2 + rtol 5e-2 (Linear.Cuda.cpp:121-129)."""
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
MX4_DEFAULT = 2          # largest M the packed-nibble kind::mxf4 kernels take by default (decode_mx4.cu)


def _check(y_t, yf_ref, tol=TOL):
    y = O.bf16_bits_to_f32(G.bits_of(y_t)).reshape(yf_ref.shape)
    err = H.rel_err_rowabs(y, yf_ref)
    assert err <= tol, f"row-abs relative error {err:.4g} > {tol}"
    assert np.all(np.abs(y - yf_ref) <= 5e-2 + 5e-2 * np.abs(yf_ref))


def _run(policy, N, K, M, bias=False, seed=0, xbits=None):
    w = H.xavier_weights_bf16(N, K, seed=1234 + seed)
    x = H.activations_bf16(M, K, seed=99 + seed) if xbits is None else xbits
    b = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(N))) if bias else None
    wd = G.bf16_tensor(w, "cuda")
    if isinstance(policy, PerChannelFp8):
        q, s = quantize_fp8_per_channel(wd)
        _, yf = O.linear_forward_fp8(x, G.u8(q), G.f32(s), b)
    else:
        g = policy.kQuantizationGroupSize
        q, s = quantize_fp4_per_group(wd, g)
        _, yf = O.linear_forward_fp4(x, G.u8(q), G.f32(s), g, b)
    xd = G.bf16_tensor(x, "cuda")
    bd = None if b is None else G.bf16_tensor(b, "cuda")
    y = linear_forward(xd, q, s, policy, bd)
    torch.cuda.synchronize()
    return y, yf, (xd, q, s, bd)


POLICIES = [PerChannelFp8(), PerGroupFp4(128), PerGroupFp4(64)]


@pytest.mark.parametrize("policy", POLICIES, ids=["fp8", "fp4g128", "fp4g64"])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 7, 8, 9, 16])
@pytest.mark.parametrize("N,K", [(32, 128), (256, 512), (40, 1024), (3840, 4096)])
def test_decode_matches_oracle(policy, M, N, K):
    y, yf, _ = _run(policy, N, K, M, seed=M)
    _check(y, yf)
    name = _lib.last_kernel()
    if K % 128 == 0 and isinstance(policy, PerGroupFp4) and policy.kQuantizationGroupSize == 128 and M <= MX4_DEFAULT:
        assert name.startswith("decode_mx4_kernel"), name           # packed nibbles -> tcgen05 kind::mxf4
    elif K % 128 == 0 and not (isinstance(policy, PerGroupFp4) and policy.kQuantizationGroupSize == 64):
        assert name.startswith("decode_tc_kernel"), name            # TMA + tcgen05 primary path
    else:
        assert name.startswith(("gemv_flat_kernel", "gemv_mma_kernel")), name


@pytest.fixture
def no_mx4():
    """Routes FP4 M <= 8 decode to decode_tc.cu (kind::f8f6f4 over unpacked nibbles) instead of decode_mx4.cu."""
    _lib.set_option("decode_mx4_max_m", 0)
    yield
    _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)


@pytest.mark.parametrize("M", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("N,K", [(256, 512), (3840, 4096), (200, 1152)])
def test_fp4_small_m_through_decode_tc_matches_oracle(M, N, K, no_mx4):
    y, yf, _ = _run(PerGroupFp4(128), N, K, M, seed=M)
    _check(y, yf)
    assert _lib.last_kernel().startswith("decode_tc_kernel")


@pytest.mark.parametrize("M", [1, 2])
@pytest.mark.parametrize("N,K,bias", [(128, 128, False), (128, 256, False), (256, 512, True), (200, 1152, True),
                                      (3840, 4096, False), (3840, 15360, False), (30720, 3840, False), (100, 3968, True)])
def test_fp4_packed_mxf4_decode_matches_oracle(M, N, K, bias):
    """decode_mx4.cu: ragged N, K with an odd number of groups / a half-filled 256-k row / a partial unit, split-K
    shapes, bias; and agreement with the independent decode_tc.cu path."""
    try:
        y, yf, (xd, q, s, bd) = _run(PerGroupFp4(128), N, K, M, bias=bias, seed=10 + M)
        assert _lib.last_kernel().startswith("decode_mx4_kernel"), _lib.last_kernel()
        _check(y, yf)
        y2 = linear_forward(xd, q, s, PerGroupFp4(128), bd)
        torch.cuda.synchronize()
        assert torch.equal(y, y2)                                     # deterministic
        _lib.set_option("decode_mx4_max_m", 0)
        y3 = linear_forward(xd, q, s, PerGroupFp4(128), bd)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)
    assert H.rel_err_rowabs(y.float().cpu().numpy(), y3.float().cpu().numpy()) <= 1e-2


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
@pytest.mark.parametrize("N,K,M", [(3840, 4096, 1), (14336, 4096, 16), (3840, 15360, 5), (200, 1024, 2), (28672, 8192, 12)])
def test_stream_k_decomposition_matches_item_split(policy, N, K, M):
    """The stream-K decomposition (used automatically for unbalanced multi-wave shapes at M > 8) cuts tiles at
    arbitrary unit boundaries; results must match the oracle, be deterministic, and agree with the item split."""
    L = _lib.lib()
    _lib.set_option("decode_streamk", 1)
    try:
        y1, yf, (xd, q, s, bd) = _run(policy, N, K, M, seed=3)
        _check(y1, yf)
        y2 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
        assert torch.equal(y1, y2)
        _lib.set_option("decode_streamk", 0)
        y3 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("decode_streamk", -1)
    assert H.rel_err_rowabs(y1.float().cpu().numpy(), y3.float().cpu().numpy()) <= 1e-2


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
@pytest.mark.parametrize("N,K,M,bias", [(256, 512, 9, False), (128, 128, 16, True), (200, 1152, 11, True), (3840, 4096, 16, False),
                                        (14336, 4096, 16, False), (4096, 14336, 13, False), (3840, 15360, 10, True),
                                        (28672, 8192, 12, False), (100, 3968, 16, True)])
def test_presplit_activations_match_converter_warps_bit_for_bit(policy, N, K, M, bias):
    """M > 8: the activation split runs once in act_presplit_kernel and the decode kernel bulk-copies the planes.
    Same arithmetic, same bytes as the in-kernel converter warps: outputs must be bit-identical (and match the
    oracle); covers ragged N, an odd number of groups (partial last unit), split-K clusters, two-wave shapes."""
    L = _lib.lib()
    try:
        _lib.set_option("decode_presplit", 1)
        y1, yf, (xd, q, s, bd) = _run(policy, N, K, M, bias=bias, seed=20 + M)
        assert _lib.last_kernel().startswith("decode_tc_kernel"), _lib.last_kernel()
        _check(y1, yf)
        _lib.reset_launch_count()
        y2 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
        assert _lib.launch_count() == 2                                # pre-pass + decode
        assert torch.equal(y1, y2)                                     # deterministic
        _lib.set_option("decode_presplit", 0)
        _lib.reset_launch_count()
        y3 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
        assert _lib.launch_count() == 1
    finally:
        _lib.set_option("decode_presplit", 1)
    assert torch.equal(y1, y3)


@pytest.fixture
def mma_sync_only():
    """Routes decode to the mma.sync kernels (the path for shapes the tcgen05 kernels do not take)."""
    _lib.set_option("decode_tc", 0)
    _lib.set_option("decode_mx4_max_m", 0)
    yield
    _lib.set_option("decode_tc", 1)
    _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)


@pytest.mark.parametrize("policy", POLICIES, ids=["fp8", "fp4g128", "fp4g64"])
@pytest.mark.parametrize("M", [1, 3, 8, 16])
@pytest.mark.parametrize("N,K", [(256, 512), (3840, 4096)])
def test_decode_mma_sync_path_matches_oracle(policy, M, N, K, mma_sync_only):
    y, yf, _ = _run(policy, N, K, M, seed=M)
    _check(y, yf)
    assert _lib.last_kernel().startswith(("gemv_flat_kernel", "gemv_mma_kernel"))


# (digit planes per MMA, activation split: 2 = converter warps of every CTA, 1 = cooperative in-launch image, 0 = pre-pass kernel)
MX8_MODES = [(3, 1), (3, 2), (2, 1), (1, 2), (3, 0)]
MX8_IDS = ["triple-coop", "triple-inkernel", "pair-coop", "single-inkernel", "triple-prepass"]


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
@pytest.mark.parametrize("N,K,M", [(3840, 4096, 1), (14336, 4096, 16), (3840, 15360, 5), (200, 1024, 2)])
def test_decode_is_deterministic_and_paths_agree(policy, N, K, M):
    """Stream-K fix-up adds partials in a fixed order: two forwards give identical bits
    (Linear.Cuda.cpp:744 compares forwards with EXPECT_EQ); and the tcgen05 path agrees with the
    independent mma.sync path to BF16 rounding."""
    y1, yf, (xd, q, s, bd) = _run(policy, N, K, M, seed=7)
    assert _lib.last_kernel().startswith(("decode_tc_kernel", "decode_mx4_kernel"))
    for _ in range(3):
        y2 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
        assert torch.equal(y1, y2)
    _lib.set_option("decode_tc", 0)
    _lib.set_option("decode_mx4_max_m", 0)
    try:
        y3 = linear_forward(xd, q, s, policy, bd)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("decode_tc", 1)
        _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)
    a = y1.float().cpu().numpy(); b = y3.float().cpu().numpy()
    assert H.rel_err_rowabs(a, b) <= 1e-2


@pytest.mark.parametrize("policy", POLICIES, ids=["fp8", "fp4g128", "fp4g64"])
def test_reference_test_shape_64_to_32(policy):
    """Linear.Cuda.cpp:265/310 dims (64 -> 32), shapes {2,4,64} and {1,1,64}, with bias."""
    if isinstance(policy, PerGroupFp4) and policy.kQuantizationGroupSize == 128:
        pytest.skip("K=64 < group 128")
    for M in (1, 8):
        y, yf, _ = _run(policy, 32, 64, M, bias=True)
        _check(y, yf)


@pytest.mark.parametrize("policy", POLICIES, ids=["fp8", "fp4g128", "fp4g64"])
def test_ragged_rows_and_bias(policy):
    # N not a multiple of the 16-row MMA tile, bias on
    for N in (1, 15, 17, 250):
        y, yf, _ = _run(policy, N, 256, 3, bias=True, seed=N)
        _check(y, yf)


def test_generic_fallback_for_odd_k():
    # K % 64 != 0 -> one-warp-per-row fallback kernel
    y, yf, _ = _run(PerChannelFp8(), 48, 72, 2)
    _check(y, yf)
    assert "generic" in _lib.last_kernel()


def test_fifteen_decade_rows_fp4():
    """Linear.Cuda.cpp:773-875 fixture: M=16 rows of magnitude 1e-8..1e7, K=512, N=256, FP4 g=128;
    the batched result must equal the per-row decode result within 1e-1*row_absmax (we hold 1e-2)."""
    wb = O.ref_weight_blob(256, 512)
    q, s = quantize_fp4_per_group(G.bf16_tensor(wb, "cuda"), 128)
    xb = O.f32_to_bf16_bits(O.ref_magnitude_rows(16, 512))
    xd = G.bf16_tensor(xb, "cuda")
    pol = PerGroupFp4(128)
    y16 = linear_forward(xd, q, s, pol)
    _, yf = O.linear_forward_fp4(xb, G.u8(q), G.f32(s), 128)
    _check(y16, yf)
    for m in range(16):
        y1 = linear_forward(xd[m:m + 1], q, s, pol)
        a = y1.float().cpu().numpy().reshape(-1); b = y16[m].float().cpu().numpy()
        row_absmax = np.max(np.abs(a))
        assert np.all(np.abs(a - b) <= 1e-2 * row_absmax + 1e-30)


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
def test_special_activation_values(policy):
    K, N = 512, 64
    x = H.activations_bf16(4, K)
    xf = O.bf16_bits_to_f32(x).copy()
    xf[0, :] = 0.0                       # all-zero token
    xf[1, :] *= 1e-30                    # tiny token
    xf[2, :] *= 3e30                     # huge token
    xf[3, 5] = 1e4                       # one outlier
    y, yf, _ = _run(policy, N, K, 4, xbits=O.f32_to_bf16_bits(xf))
    _check(y, yf)


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
@pytest.mark.parametrize("M", [17, 40, 128])
def test_batched_slots_large_m(policy, M):
    y, yf, _ = _run(policy, 256, 512, M, bias=True)
    _check(y, yf)


@pytest.mark.skipif(not O.ref_lib_path().exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("policy", POLICIES, ids=["fp8", "fp4g128", "fp4g64"])
def test_m1_matches_reference_matvec_kernel(policy):
    """Same inputs through Mila's own matvec kernel (recompiled for sm_100a): both are FP32-accumulate
    paths over identical weights, so they agree to BF16 rounding of nearly equal sums."""
    for (N, K) in [(256, 512), (3840, 4096), (4096, 14336)]:
        y, yf, (xd, q, s, bd) = _run(policy, N, K, 1, bias=True)
        g = 0 if isinstance(policy, PerChannelFp8) else policy.kQuantizationGroupSize
        yr = G.ref_matvec(xd, q, s, g, bd)
        a = y.float().cpu().numpy().reshape(-1); b = yr.float().cpu().numpy()
        assert H.rel_err_rowabs(a[None], b[None]) <= 1e-2
        # and the generic device kernel agrees too (independent second implementation)
        yg = G.generic_gemv(xd, q, s, g, bd).float().cpu().numpy().reshape(-1)
        assert H.rel_err_rowabs(a[None], yg[None]) <= 1e-2


@pytest.mark.parametrize("name,policy,N,K", [
    ("llama8b_gate_fp8", PerChannelFp8(), 14336, 4096),
    ("llama8b_down_fp8", PerChannelFp8(), 4096, 14336),
    ("gemma_qkv_fp4", PerGroupFp4(128), 8192, 3840),
    ("gemma_o_fp4", PerGroupFp4(128), 3840, 4096),
    ("gemma_down_fp4", PerGroupFp4(128), 3840, 15360),
])
@pytest.mark.parametrize("M", [1, 16])
def test_full_size_config_shapes(name, policy, N, K, M):
    """BASELINE.json config shapes at full size vs the oracle (seconds on CPU at M <= 16)."""
    y, yf, _ = _run(policy, N, K, M)
    _check(y, yf)


def test_linearity_property_full_size():
    """Size-independent property at a full-size shape: f(a*x) == a*f(x) for a power-of-two a (exact in
    every stage of the kernel), and f(x) for x = e_k picks out column k of the dequantised weights."""
    N, K = 14336, 4096
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = quantize_fp8_per_channel(w)
    x = torch.randn((1, K), device="cuda").to(torch.bfloat16)
    pol = PerChannelFp8()
    y1 = linear_forward(x, q, s, pol).clone()
    y2 = linear_forward(x * 4.0, q, s, pol)
    assert torch.equal(y1.float() * 4.0, y2.float())
    e = torch.zeros((1, K), device="cuda", dtype=torch.bfloat16); e[0, 1234] = 1.0
    col = linear_forward(e, q, s, pol).float().cpu().numpy().reshape(-1)
    wf = O.dequant_fp8(G.u8(q[:, 1234:1235].contiguous()), G.f32(s)).reshape(-1)
    np.testing.assert_array_equal(col, O.bf16_bits_to_f32(O.f32_to_bf16_bits(wf)))
