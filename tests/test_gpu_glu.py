"""GPU parity of the gate|up Linear with the gated activation fused into its epilogue
and of the stand-alone GeGLU / SwiGLU kernels.

Pins: (1) the stand-alone kernels equal the reference's own kernels (oracle/_ref, compiled unmodified from
Activations/{Geglu,Swiglu}/Kernels) bit for bit; (2) the fused call equals the two-step sequence it replaces
(our Linear, then the reference's activation kernel) bit for bit; (3) both are within tolerance of the CPU
oracle (dequantise-then-FP32-GEMM, BF16 rounding, FP32 activation)."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import (GLU_GEGLU_TANH, GLU_SWIGLU, PerChannelFp8, PerGroupFp4, glu_forward, linear_forward,
                              linear_glu_forward, quantize_fp4_per_group, quantize_fp8_per_channel)
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KINDS = [GLU_GEGLU_TANH, GLU_SWIGLU]
KIDS = ["geglu", "swiglu"]


def _ref_glu(x: torch.Tensor, kind: int) -> torch.Tensor:
    R = O.ref_lib()
    tokens, twoH = x.shape
    y = torch.empty((tokens, twoH // 2), dtype=torch.bfloat16, device="cuda")
    fn = R.milaref_geglu_forward_bf16 if kind == GLU_GEGLU_TANH else R.milaref_swiglu_forward_bf16
    rc = fn(G.p(y), G.p(x), y.numel(), twoH // 2, ctypes.c_void_p(G.stream()))
    torch.cuda.synchronize(); assert rc == 0
    return y


needs_ref = pytest.mark.skipif(not O.ref_lib_path().exists(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("tokens,H", [(1, 128), (3, 1024), (16, 15360), (7, 136)])
def test_standalone_glu_equals_reference_kernel(kind, tokens, H):
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    x = (torch.randn((tokens, 2 * H), device="cuda", generator=g) * 3.0).to(torch.bfloat16)
    x[0, :8] = torch.tensor([0.0, -0.0, 1e-30, -1e-30, 50.0, -50.0, 1e4, -1e4], device="cuda").to(torch.bfloat16)
    if kind == GLU_SWIGLU and H % 8 != 0:
        pytest.skip("the reference SwiGLU kernel is uint4-vectorised (half_width % 8 == 0)")
    y = glu_forward(x, kind)
    torch.cuda.synchronize()
    assert torch.equal(y, _ref_glu(x, kind))
    # CPU oracle within one BF16 ulp (libm vs device transcendental functions)
    yo = O.bf16_bits_to_f32(O.glu_forward_bf16(G.bits_of(x), kind))
    yf = y.float().cpu().numpy()
    assert np.all(np.abs(yf - yo) <= 2.0 ** -7 * np.abs(yo) + 1e-30)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
@pytest.mark.parametrize("Hh,K,M", [(14336, 4096, 1), (15360, 3840, 2), (14336, 4096, 5), (15360, 3840, 16), (14336, 1152, 3)])
def test_fused_gate_up_glu_equals_two_step_sequence(kind, policy, Hh, K, M):
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    w = (torch.randn((2 * Hh, K), device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    x = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    q, s = quantize_fp8_per_channel(w) if isinstance(policy, PerChannelFp8) else quantize_fp4_per_group(w, 128)
    y = linear_glu_forward(x, q, s, policy, kind)
    torch.cuda.synchronize()
    name = _lib.last_kernel()
    assert name.startswith(("decode_tc_kernel", "decode_mx4_kernel")), name          # one launch: fused
    # the two-step Linear must accumulate whole rows like the fused kernel does (stream-K, which the decode
    # kernel picks for unbalanced shapes at M > 8, groups the FP32 sum differently: one-ulp BF16 differences)
    _lib.set_option("decode_streamk", 0)
    try:
        gate_up = linear_forward(x, q, s, policy)
    finally:
        _lib.set_option("decode_streamk", -1)
    two_step = _ref_glu(gate_up, kind) if O.ref_lib_path().exists() else glu_forward(gate_up, kind)
    torch.cuda.synchronize()
    assert torch.equal(y, two_step)
    # deterministic
    y2 = linear_glu_forward(x, q, s, policy, kind)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("name,policy,Hh,K,M,bias,fused", [
    ("fp8", PerChannelFp8(), 1024, 512, 1024, False, True),
    ("fp8_ragged_bias", PerChannelFp8(), 1000, 640, 1000, True, True),
    ("fp8_llama8b", PerChannelFp8(), 14336, 4096, 2048, False, True),
    ("fp8_extreme_bias", PerChannelFp8(), 1024, 512, 1024, True, True),
    ("fp8_few_tiles_k_split", PerChannelFp8(), 512, 2048, 256, False, False),
    ("fp4_hi_lo", PerGroupFp4(128), 1024, 512, 1024, False, False),
    ("fp4_sum_ragged_bias", PerGroupFp4(128), 2000, 384, 1900, True, False),
])
def test_batched_gate_up_glu_equals_two_step_sequence(kind, name, policy, Hh, K, M, bias, fused):
    """M > 32, FP8 weights: the activation in the epilogue of the batched TMA + tcgen05 kernel (CTA pairs: rank 0 projects 128
    gate rows, rank 1 the 128 up rows H below; each rank finishes half of the tokens, the halves cross over distributed shared
    memory).  Bit-identical to the batched Linear [M, 2H] followed by the reference's activation kernel — incl. H and M that
    are not multiples of the tile, and bias — deterministic, and within tolerance of an FP32 reference.  FP4 weights and the
    k-split regime keep the two-kernel sequence (measured faster there): same entry, same bits."""
    want = "prefill_tc_kernel<fp8,cta_pair,glu>" if fused else "glu_forward_bf16_kernel"
    g = torch.Generator(device="cuda"); g.manual_seed(123 + Hh)
    w = (torch.randn((2 * Hh, K), device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    x = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(2 * Hh, device="cuda", generator=g) * 0.1).to(torch.bfloat16) if bias else None
    if name.endswith("extreme_bias"):
        # gate projections far outside the activation's comfortable range: 1 + e^-g overflows FP32 below g = -88.7 and leaves
        # the reciprocal's fast path below g = -87.3 (the fused epilogue re-evaluates those with the reference's own expression)
        vals = torch.tensor([-300.0, -200.0, -100.0, -90.0, -89.0, -88.5, -88.0, -87.5, -87.0, -86.0, -50.0, 50.0, 88.0, 100.0, 300.0, 0.0],
                            device="cuda").to(torch.bfloat16)
        b[:256] = vals.repeat(16); b[Hh:Hh + 64] = vals.repeat(4)
    q, s = quantize_fp8_per_channel(w) if isinstance(policy, PerChannelFp8) else quantize_fp4_per_group(w, 128)
    before = _lib.launch_count()
    y = linear_glu_forward(x, q, s, policy, kind, b).clone()
    torch.cuda.synchronize()
    assert _lib.last_kernel().startswith(want), _lib.last_kernel()
    assert _lib.launch_count() - before == (2 if fused else 3)    # fused: activation pre-pass + ONE GEMM (no [M, 2H] round trip)
    gate_up = linear_forward(x, q, s, policy, b)
    assert _lib.last_kernel().startswith("prefill_tc_kernel") and not _lib.last_kernel().endswith("glu>")
    two_step = _ref_glu(gate_up, kind) if O.ref_lib_path().exists() else glu_forward(gate_up, kind)
    torch.cuda.synchronize()
    assert torch.equal(y, two_step)
    y2 = linear_glu_forward(x, q, s, policy, kind, b); torch.cuda.synchronize()
    assert torch.equal(y, y2)
    # the unfused route (option off) gives the same bits
    _lib.set_option("prefill_glu", 0)
    try:
        y3 = linear_glu_forward(x, q, s, policy, kind, b); torch.cuda.synchronize()
        assert _lib.last_kernel().startswith("glu_forward_bf16_kernel")
    finally:
        _lib.set_option("prefill_glu", 1)
    assert torch.equal(y, y3)
    if Hh <= 2048 and not name.endswith("extreme_bias"):
        # FP32 reference of the whole thing (projections rounded to BF16 as the Linear stores them)
        import test_gpu_prefill as TP
        ref_lin = TP._torch_ref(policy, x, q, s, b).to(torch.bfloat16).float()
        gate, up = ref_lin[:, :Hh], ref_lin[:, Hh:]
        act = torch.nn.functional.gelu(gate, approximate="tanh") if kind == GLU_GEGLU_TANH else torch.nn.functional.silu(gate)
        ref = act * up
        row_abs = ref.abs().amax(dim=1, keepdim=True)
        assert bool(((y.float() - ref).abs() <= 3e-2 * row_abs + 1e-6).all())


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128), PerGroupFp4(64)], ids=["fp8", "fp4g128", "fp4g64"])
def test_glu_matches_cpu_oracle_and_unfused_fallback(kind, policy):
    """Small / ineligible shapes take the Linear + stand-alone activation route (few logical tiles, g = 64, bias,
    M > 16); the fused big shape is checked against the CPU oracle on its first rows."""
    for (Hh, K, M, bias) in [(256, 512, 3, True), (128, 256, 40, False), (384, 1024, 16, True)]:
        w = H.xavier_weights_bf16(2 * Hh, K, seed=Hh)
        x = H.activations_bf16(M, K, seed=M)
        b = O.f32_to_bf16_bits(O.ref_bias_value(np.arange(2 * Hh))) if bias else None
        wd = G.bf16_tensor(w, "cuda")
        if isinstance(policy, PerChannelFp8):
            q, s = quantize_fp8_per_channel(wd); lin_bits, _ = O.linear_forward_fp8(x, G.u8(q), G.f32(s), b)
        else:
            gsz = policy.kQuantizationGroupSize
            q, s = quantize_fp4_per_group(wd, gsz); lin_bits, _ = O.linear_forward_fp4(x, G.u8(q), G.f32(s), gsz, b)
        ref = O.bf16_bits_to_f32(O.glu_forward_bf16(lin_bits, kind))
        bd = None if b is None else G.bf16_tensor(b, "cuda")
        y = linear_glu_forward(G.bf16_tensor(x, "cuda"), q, s, policy, kind, bd).float().cpu().numpy()
        # the oracle's BF16 projections can differ from the device's by one ulp (FP32 summation order), which
        # the activation carries through: reference BF16 budget
        assert np.all(np.abs(y - ref) <= 5e-2 + 5e-2 * np.abs(ref)), (Hh, K, M)
        assert H.rel_err_rowabs(y, ref) <= 2e-2, (Hh, K, M)


def test_glu_argument_errors():
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.milab200_fp4a16_gemm_glu(one, one, one, one, one, None, 1, 128, 128, 32, 1, None) == _lib.E_UNSUPPORTED_GROUP
    assert L.milab200_w8a16_gemm_glu(one, one, one, one, one, None, 1, 128, 128, 7, None) == _lib.E_INVALID_ARGUMENT
    assert L.milab200_geglu_forward_bf16(None, one, 128, 128, None) == _lib.E_INVALID_ARGUMENT
