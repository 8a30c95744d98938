"""GPU parity: load-time quantizers through the C-ABI, bit-exact against (a) the CPU oracle and
(b) the reference's own kernels compiled unmodified into oracle/_ref."""
import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import quantize_fp4_per_group, quantize_fp8_per_channel
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [(32, 64), (48, 256), (17, 1024), (256, 512), (64, 3840), (33, 4096), (16, 14336), (8, 28672)]


def _have_ref():
    return O.ref_lib_path().exists()


@pytest.mark.parametrize("N,K", SHAPES)
def test_fp8_bit_exact_vs_oracle(N, K):
    w = H.xavier_weights_bf16(N, K, seed=1234 + N)
    q, s = quantize_fp8_per_channel(G.bf16_tensor(w, "cpu", pinned=True))
    torch.cuda.synchronize()
    qo, so = O.quantize_fp8_per_channel(w)
    np.testing.assert_array_equal(G.f32(s).view(np.uint32), so.view(np.uint32))
    np.testing.assert_array_equal(G.u8(q), qo)


@pytest.mark.parametrize("g", [64, 128])
@pytest.mark.parametrize("N,K", [s for s in SHAPES if s[1] % 128 == 0] + [(32, 64)])
def test_fp4_bit_exact_vs_oracle(N, K, g):
    if K % g:
        pytest.skip("K % g")
    w = H.xavier_weights_bf16(N, K, seed=1234 + N)
    q, s = quantize_fp4_per_group(G.bf16_tensor(w, "cpu", pinned=True), g)
    torch.cuda.synchronize()
    qo, so = O.quantize_fp4_per_group(w, g)
    np.testing.assert_array_equal(G.f32(s).view(np.uint32), so.view(np.uint32))
    np.testing.assert_array_equal(G.u8(q), qo)


def test_adversarial_pack_vs_oracle_and_reference():
    w = H.adversarial_weights_bf16(16, 512, group=128)
    q8, s8 = quantize_fp8_per_channel(G.bf16_tensor(w, "cpu"))          # pageable host source
    qo, so = O.quantize_fp8_per_channel(w)
    np.testing.assert_array_equal(G.f32(s8).view(np.uint32), so.view(np.uint32))
    np.testing.assert_array_equal(G.u8(q8), qo)
    for g in (64, 128):
        q4, s4 = quantize_fp4_per_group(G.bf16_tensor(w, "cpu"), g)
        qo4, so4 = O.quantize_fp4_per_group(w, g)
        np.testing.assert_array_equal(G.f32(s4).view(np.uint32), so4.view(np.uint32))
        np.testing.assert_array_equal(G.u8(q4), qo4)
    if _have_ref():
        qr, sr = G.ref_quantize_fp8(w)
        np.testing.assert_array_equal(G.u8(q8), qr)
        np.testing.assert_array_equal(G.f32(s8).view(np.uint32), sr.view(np.uint32))
        for g in (64, 128):
            q4, s4 = quantize_fp4_per_group(G.bf16_tensor(w, "cpu"), g)
            qr4, sr4 = G.ref_quantize_fp4(w, g)
            np.testing.assert_array_equal(G.u8(q4), qr4)
            np.testing.assert_array_equal(G.f32(s4).view(np.uint32), sr4.view(np.uint32))


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("N,K", [(256, 512), (96, 4096), (40, 14336)])
def test_bit_exact_vs_reference_kernels(N, K):
    """Packed bytes and scales equal Mila's own kernels' output on the same blob."""
    w = H.xavier_weights_bf16(N, K, seed=77)
    q8, s8 = quantize_fp8_per_channel(G.bf16_tensor(w, "cpu", pinned=True))
    qr, sr = G.ref_quantize_fp8(w)
    np.testing.assert_array_equal(G.u8(q8), qr)
    np.testing.assert_array_equal(G.f32(s8).view(np.uint32), sr.view(np.uint32))
    for g in (64, 128):
        q4, s4 = quantize_fp4_per_group(G.bf16_tensor(w, "cpu", pinned=True), g)
        qr4, sr4 = G.ref_quantize_fp4(w, g)
        np.testing.assert_array_equal(G.u8(q4), qr4)
        np.testing.assert_array_equal(G.f32(s4).view(np.uint32), sr4.view(np.uint32))
    # the reference test fixture blob (Linear.Cuda.cpp:666-677)
    wb = O.ref_weight_blob(256, 512)
    q4, s4 = quantize_fp4_per_group(G.bf16_tensor(wb, "cpu"), 128)
    qr4, sr4 = G.ref_quantize_fp4(wb, 128)
    np.testing.assert_array_equal(G.u8(q4), qr4)
    np.testing.assert_array_equal(G.f32(s4), sr4)


def test_device_source_entry_equals_host_source_entry():
    w = H.xavier_weights_bf16(64, 1024, seed=5)
    wh = G.bf16_tensor(w, "cpu", pinned=True)
    wd = G.bf16_tensor(w, "cuda")
    a, sa = quantize_fp8_per_channel(wh); b, sb = quantize_fp8_per_channel(wd)
    assert torch.equal(a, b) and torch.equal(sa, sb)
    a, sa = quantize_fp4_per_group(wh, 128); b, sb = quantize_fp4_per_group(wd, 128)
    assert torch.equal(a, b) and torch.equal(sa, sb)


def test_full_size_quantize_properties():
    """BASELINE config shapes at full size: size-independent properties instead of the scalar oracle —
    (1) idempotence: quantize(dequant(quantize(w))) == quantize(w) bytes;
    (2) a row-sample equals the oracle bit for bit;  (3) scale == absmax/C computed by torch."""
    torch.manual_seed(1234)
    for (N, K) in [(14336, 4096), (3840, 15360)]:
        w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
        q8, s8 = quantize_fp8_per_channel(w)
        q4, s4 = quantize_fp4_per_group(w, 128)
        torch.cuda.synchronize()
        absmax = w.float().abs().amax(dim=1)
        assert torch.equal(s8, torch.div(absmax, torch.full_like(absmax, 448.0)))   # true division, not x*(1/448)
        gmax = w.float().abs().view(N, K // 128, 128).amax(dim=2)
        assert torch.equal(s4, torch.div(gmax, torch.full_like(gmax, 6.0)))
        rows = [0, 1, N // 2, N - 1]
        wb = G.bits_of(w[rows])
        qo, so = O.quantize_fp8_per_channel(wb)
        np.testing.assert_array_equal(G.u8(q8[rows]), qo)
        qo4, so4 = O.quantize_fp4_per_group(wb, 128)
        np.testing.assert_array_equal(G.u8(q4[rows]), qo4)
        # idempotence through the dequantised BF16-representable FP4 grid: lut*scale rounds to bf16,
        # so compare nibbles after re-quantising the *exact* products where they are representable
        lut = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6], device="cuda")
        nib = torch.stack([q4 & 0xF, q4 >> 4], dim=-1).view(N, K)
        deq = lut[(nib & 7).long()] * torch.where((nib & 8) > 0, -1.0, 1.0) * s4.repeat_interleave(128, dim=1)
        exact = deq.to(torch.bfloat16).float() == deq
        q4b, s4b = quantize_fp4_per_group(deq.to(torch.bfloat16), 128)
        nibb = torch.stack([q4b & 0xF, q4b >> 4], dim=-1).view(N, K)
        grp_exact = exact.view(N, K // 128, 128).all(dim=2).repeat_interleave(128, dim=1)
        same = (nibb == nib) | ((nib & 7) == 0)
        assert bool(same[grp_exact].all())


def test_error_codes_on_device():
    w = G.bf16_tensor(H.xavier_weights_bf16(8, 256), "cuda")
    with pytest.raises(_lib.MilaB200Error):
        quantize_fp4_per_group(w, 32)
    with pytest.raises(_lib.InvalidArgument):
        quantize_fp4_per_group(w[:, :192].contiguous(), 128)
