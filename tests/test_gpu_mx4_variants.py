"""GPU parity of the opt-in kind::mxf4 decode variants (4-token converter variant, 8-token pre-split variant with
hardware digit recombination).  They lose to decode_tc.cu on every routed shape, so they are compiled
only into the diagnostics build:

    make -C mila_b200/csrc diag && MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_diag.so python -m pytest tests/test_gpu_mx4_variants.py -m gpu

With the product library loaded (the default) this module skips."""
import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import PerGroupFp4, linear_forward, quantize_fp4_per_group
from oracle import oracle as O
from test_gpu_gemv import _check, _run

pytestmark = pytest.mark.gpu
MX4_DEFAULT = 2


@pytest.fixture(scope="module", autouse=True)
def _needs_diag_build():
    if _lib.lib().milab200_set_option(b"mx8_pair", 3) != 0:
        pytest.skip("product build: the kind::mxf4 multi-token variants live in the diagnostics build (MILAB200_LIB)")
    yield


@pytest.mark.parametrize("M", [3, 4])
@pytest.mark.parametrize("N,K,bias", [(128, 128, False), (128, 256, False), (256, 512, True), (200, 1152, True),
                                      (3840, 4096, False), (3840, 15360, False), (30720, 3840, False), (100, 3968, True)])
def test_fp4_packed_mxf4_four_token_variant_matches_oracle(M, N, K, bias):
    """decode_mx4.cu: ragged N, K with an odd number of groups / a half-filled 256-k row / a partial unit, split-K
    shapes, bias; and agreement with the independent decode_tc.cu path."""
    _lib.set_option("decode_mx4_max_m", 4)                        # the 4-token variant too (default routes M <= 2)
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



# (digit planes per MMA, activation split: 2 = converter warps of every CTA, 1 = cooperative in-launch image, 0 = pre-pass kernel)
MX8_MODES = [(3, 1), (3, 2), (2, 1), (1, 2), (3, 0)]
MX8_IDS = ["triple-coop", "triple-inkernel", "pair-coop", "single-inkernel", "triple-prepass"]


@pytest.fixture(params=MX8_MODES, ids=MX8_IDS)
def mx8_pair(request):
    """8-token kind::mxf4 variant: three (default) / two / one digit planes per MMA with per-column scale factors;
    activations split cooperatively inside the launch (default) / by every CTA's converter warps / by a pre-pass kernel."""
    _lib.set_option("mx8_pair", request.param[0])
    _lib.set_option("mx8_coop", request.param[1])
    yield request.param
    _lib.set_option("mx8_pair", 3)
    _lib.set_option("mx8_coop", 1)


_MX8_SMALL = [(128, 128, False), (128, 256, False), (256, 512, True), (200, 1152, True), (100, 3968, True), (3840, 4096, False)]
_MX8_BIG = [(3840, 15360, False), (30720, 3840, False), (8192, 28672, False)]       # split-K clusters, two waves, long K
# the default mode on every shape (every M on the small ones); the other modes on a ragged and a split-K shape
_MX8_CASES = [(M, N, K, b, MX8_MODES[0]) for M in (3, 4, 5, 8) for (N, K, b) in _MX8_SMALL] + \
             [(M, N, K, b, MX8_MODES[0]) for M in (4, 8) for (N, K, b) in _MX8_BIG] + \
             [(M, N, K, b, mode) for mode in MX8_MODES[1:] for M in (3, 8)
              for (N, K, b) in [(200, 1152, True), (3840, 15360, False)]]


@pytest.mark.parametrize("M,N,K,bias,mx8_pair", _MX8_CASES, indirect=["mx8_pair"],
                         ids=[f"{MX8_IDS[MX8_MODES.index(c[4])]}-{c[1]}x{c[2]}-m{c[0]}" for c in _MX8_CASES])
def test_fp4_presplit_mxf4_decode_matches_oracle(M, N, K, bias, mx8_pair):
    """decode_mx4.cu, 8-token variant (opt-in for M = 3..8): activations pre-split into six signed base-8 digit
    planes by act_presplit_mx4_kernel, bulk-copied by the producer.  Ragged N, odd group counts, half-filled 256-k
    rows, partial units, split-K clusters, bias; deterministic; agrees with the independent decode_tc.cu path."""
    try:
        _lib.set_option("decode_mx4_max_m", 8)
        y, yf, (xd, q, s, bd) = _run(PerGroupFp4(128), N, K, M, bias=bias, seed=30 + M)
        assert _lib.last_kernel().startswith("decode_mx4_kernel<fp4g128,packed,t8"), _lib.last_kernel()
        _check(y, yf)
        y2 = linear_forward(xd, q, s, PerGroupFp4(128), bd)
        torch.cuda.synchronize()
        assert torch.equal(y, y2)
        _lib.set_option("decode_mx4_max_m", 0)
        y3 = linear_forward(xd, q, s, PerGroupFp4(128), bd)
        torch.cuda.synchronize()
        assert _lib.last_kernel().startswith("decode_tc_kernel")
    finally:
        _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)
    assert H.rel_err_rowabs(y.float().cpu().numpy(), y3.float().cpu().numpy()) <= 1e-2


def test_fp4_presplit_mxf4_exact_on_power_of_two_block(mx8_pair):
    """Activations whose 128-k blocks hold one non-zero power of two pick out single weight columns: the six-plane
    digit split and the base-8 recombination must then be exact (output == bf16(w * x))."""
    N, K, M = 256, 1024, 8
    w = H.xavier_weights_bf16(N, K, seed=77)
    q, s = quantize_fp4_per_group(G.bf16_tensor(w, "cuda"), 128)
    x = np.zeros((M, K), dtype=np.float32)
    rng = np.random.default_rng(5)
    for m in range(M):
        for kb in range(K // 128):
            x[m, kb * 128 + rng.integers(0, 128)] = np.float32(2.0) ** rng.integers(-20, 20) * rng.choice([-1.0, 1.0])
    xb = O.f32_to_bf16_bits(x)
    _, yf = O.linear_forward_fp4(xb, G.u8(q), G.f32(s), 128, None)
    try:
        _lib.set_option("decode_mx4_max_m", 8)
        y = linear_forward(G.bf16_tensor(xb, "cuda"), q, s, PerGroupFp4(128), None)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("decode_mx4_max_m", MX4_DEFAULT)
    assert _lib.last_kernel().startswith("decode_mx4_kernel<fp4g128,packed,t8"), _lib.last_kernel()
    got = O.bf16_bits_to_f32(G.bits_of(y)).reshape(yf.shape)
    want = O.bf16_bits_to_f32(O.f32_to_bf16_bits(yf.astype(np.float32))).reshape(yf.shape)
    assert H.rel_err_rowabs(got, want) <= 2.0 ** -7            # one BF16 ulp of slack for the FP32 summation order


