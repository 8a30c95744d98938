"""PerGroupInt4 (GPTQ-style W4A16).  CPU: the C restatement of CudaW4A16Gemm.cu:88-197 against an
independent float64 numpy evaluation of the same formula.  GPU: milab200_w4a16_gemm against that oracle and against the
reference's own cuda_w4a16_gemm compiled unmodified into oracle/_ref (symmetric and asymmetric zero points, g = 64 / 128,
M = 1..40, ragged N, K not a multiple of 128)."""
import ctypes

import numpy as np
import pytest

import parity_helpers as H
from oracle import oracle as O


def make_int4(N, K, g, asym, seed=0):
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 256, (N, K // 2), dtype=np.uint8)
    s = (rng.random((N, K // g), dtype=np.float32) * 0.02 + 0.001).astype(np.float32)
    z = rng.integers(0, 256, (N, K // g // 2), dtype=np.uint8) if asym else None
    return w, s, z


def numpy_ref(x_bits, w, s, z, bias_bits, g):
    x = O.bf16_bits_to_f32(x_bits).astype(np.float64)
    N, K2 = w.shape; K = 2 * K2
    nib = np.empty((N, K), np.float64)
    nib[:, 0::2] = w & 0xF; nib[:, 1::2] = w >> 4
    if z is None:
        zero = np.full((N, K // g), 8.0)
    else:
        zero = np.empty((N, K // g), np.float64)
        zero[:, 0::2] = z & 0xF; zero[:, 1::2] = z >> 4
    wf = (nib - np.repeat(zero, g, axis=1)) * np.repeat(s.astype(np.float64), g, axis=1)
    y = x @ wf.T
    if bias_bits is not None:
        y = y + O.bf16_bits_to_f32(bias_bits).astype(np.float64)
    return y


@pytest.mark.parametrize("g", [64, 128])
@pytest.mark.parametrize("asym", [False, True], ids=["symmetric", "zero_points"])
def test_oracle_restatement_matches_float64_formula(g, asym):
    N, K, M = 24, 512, 3
    w, s, z = make_int4(N, K, g, asym, seed=g)
    x = H.activations_bf16(M, K, seed=5)
    bias = O.f32_to_bf16_bits(np.linspace(-1, 1, N, dtype=np.float32))
    yb, yf = O.w4a16_int4_forward(x, w, s, z, bias, g)
    ref = numpy_ref(x, w, s, z, bias, g)
    assert H.rel_err_rowabs(yf, ref) <= 1e-4                 # FP32 accumulation over K = 512 vs float64
    assert np.array_equal(yb, O.f32_to_bf16_bits(yf))


def test_oracle_reproduces_the_reference_kernel_golden_bit_for_bit():
    """tests/golden/int4_ref.npz: outputs of the reference's cuda_w4a16_gemm on a B200 (tools/make_golden.py)."""
    from pathlib import Path
    f = Path(__file__).resolve().parent / "golden" / "int4_ref.npz"
    if not f.exists():
        pytest.skip("golden vectors of the reference INT4 kernel not generated yet (tools/make_golden.py on a GPU box)")
    gold = np.load(f)
    for tag in ("sym_g128", "asym_g64"):
        z = gold[f"{tag}_z"] if f"{tag}_z" in gold else None
        yb, _ = O.w4a16_int4_forward(gold[f"{tag}_x"], gold[f"{tag}_w"], gold[f"{tag}_s"], z, gold[f"{tag}_bias"], int(gold[f"{tag}_g"]))
        assert np.array_equal(yb, gold[f"{tag}_y"]), tag


# ---- GPU ---------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _gpu_run(N, K, M, g, asym, bias, seed):
    import torch
    import gpu_util as G
    from mila_b200 import _lib
    from mila_b200.linear import w4a16_forward
    w, s, z = make_int4(N, K, g, asym, seed)
    x = H.activations_bf16(M, K, seed=seed + 1)
    b = O.f32_to_bf16_bits(np.random.default_rng(seed).standard_normal(N).astype(np.float32)) if bias else None
    wd, sd = torch.from_numpy(w).cuda(), torch.from_numpy(s).cuda()
    zd = torch.from_numpy(z).cuda() if z is not None else None
    xd = G.bf16_tensor(x, "cuda"); bd = G.bf16_tensor(b, "cuda") if b is not None else None
    y = w4a16_forward(xd, wd, sd, zd, g, bd)
    torch.cuda.synchronize()
    return y, (x, w, s, z, b), (xd, wd, sd, zd, bd), _lib.last_kernel()


@gpu
@pytest.mark.parametrize("g", [64, 128])
@pytest.mark.parametrize("asym", [False, True], ids=["symmetric", "zero_points"])
@pytest.mark.parametrize("N,K,M,bias", [(256, 512, 1, False), (200, 1152, 3, True), (40, 1024, 8, True), (3840, 4096, 16, False),
                                        (512, 256, 40, True)])
def test_int4_forward_matches_oracle_and_reference_kernel(g, asym, N, K, M, bias):
    import torch
    import gpu_util as G
    if asym and (K // g) % 2:
        pytest.skip("two zero points per byte: an odd group count has no defined layout (CudaW4A16Gemm.cu:107)")
    y, (x, w, s, z, b), (xd, wd, sd, zd, bd), kernel = _gpu_run(N, K, M, g, asym, bias, seed=N + M)
    assert kernel.startswith("w4a16_int4_kernel"), kernel
    got = y.float().cpu().numpy()
    _, yf = O.w4a16_int4_forward(x, w, s, z, b, g)
    assert H.rel_err_rowabs(got, yf) <= 1e-2
    assert np.allclose(got, yf, atol=5e-2, rtol=5e-2)                  # the reference's own BF16 budget (Linear.Cuda.cpp:121-129)
    if O.ref_lib_path().exists():
        R = O.ref_lib()
        yr = torch.empty_like(y)
        rc = R.milaref_w4a16_gemm(G.p(yr), G.p(xd), G.p(wd), G.p(sd), G.p(zd), G.p(bd), M, K, N, g, ctypes.c_void_p(G.stream()))
        torch.cuda.synchronize()
        assert rc == 0
        assert H.rel_err_rowabs(got, yr.float().cpu().numpy()) <= 1e-2
    # deterministic
    from mila_b200.linear import w4a16_forward
    y2 = w4a16_forward(xd, wd, sd, zd, g, bd); torch.cuda.synchronize()
    assert torch.equal(y, y2)


@gpu
def test_int4_generic_route_and_errors():
    import torch
    from mila_b200 import _lib
    y, (x, w, s, z, b), _, kernel = _gpu_run(50, 192, 2, 64, False, True, seed=3)      # K % 128 != 0 (odd group count: symmetric only)
    assert kernel == "w4a16_int4_generic_kernel"
    _, yf = O.w4a16_int4_forward(x, w, s, z, b, 64)
    assert H.rel_err_rowabs(y.float().cpu().numpy(), yf) <= 1e-2
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.milab200_w4a16_gemm(one, one, one, one, None, None, 1, 256, 8, 32, None) == _lib.E_UNSUPPORTED_GROUP
    assert L.milab200_w4a16_gemm(one, one, one, one, None, None, 1, 192, 8, 128, None) == _lib.E_BAD_SHAPE
    assert L.milab200_w4a16_gemm(None, one, one, one, None, None, 1, 256, 8, 128, None) == _lib.E_INVALID_ARGUMENT


@gpu
def test_int4_exact_on_integer_activations():
    """Small-integer activations and power-of-two scales make every product and partial sum exact in FP32: the result
    must equal the float64 formula bit for bit after one BF16 rounding."""
    import torch
    import gpu_util as G
    from mila_b200.linear import w4a16_forward
    N, K, M, g = 64, 512, 5, 128
    rng = np.random.default_rng(9)
    w = rng.integers(0, 256, (N, K // 2), dtype=np.uint8)
    s = (2.0 ** rng.integers(-6, 2, (N, K // g))).astype(np.float32)
    z = rng.integers(0, 256, (N, K // g // 2), dtype=np.uint8)
    x = O.f32_to_bf16_bits(rng.integers(-4, 5, (M, K)).astype(np.float32))
    y = w4a16_forward(G.bf16_tensor(x, "cuda"), torch.from_numpy(w).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(z).cuda(), g)
    torch.cuda.synchronize()
    ref = numpy_ref(x, w, s, z, None, g)
    assert np.array_equal(G.bits_of(y), O.f32_to_bf16_bits(ref.astype(np.float32)))
