"""GPU parity of the FP8 tied-table token-embedding gather-dequant — bit for bit against
the reference's own kernel (Embeddings/Kernels/TokenEmbedding.Fp8.cu, compiled unmodified into oracle/_ref) and
against the CPU oracle; plus the reference's own formula pin (TokenEmbedding.Cuda.cpp:555-577: scale =
row_absmax/448, gathered value within 0.07*|w| + 0.004*scale) and the tied-table relation (the embedding row
of token v equals lm_head's dequantised weight row v)."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import PerChannelFp8, linear_forward, quantize_fp8_per_channel
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _embed(ids, q, s, decode=False):
    L = _lib.lib()
    C = q.shape[1]
    y = torch.empty((ids.numel(), C), dtype=torch.bfloat16, device="cuda")
    st = ctypes.c_void_p(G.stream())
    if decode:
        rc = L.milab200_token_embedding_decode_bf16_qfp8(G.p(y), G.p(ids), G.p(q), G.p(s), ids.numel(), C, st)
    else:
        rc = L.milab200_token_embedding_forward_bf16_qfp8(G.p(y), G.p(ids), G.p(q), G.p(s), 1, ids.numel(), C, st)
    torch.cuda.synchronize(); _lib.check(rc, "embed")
    return y


@pytest.mark.parametrize("V,C,T", [(1000, 128, 17), (5000, 3840, 64), (300, 72, 5), (262144, 3840, 33)])
def test_gather_dequant_is_bit_exact(V, C, T):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    w = (torch.randn((V, C), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    w[0] = 0                                                       # all-zero row: scale 1
    q, s = quantize_fp8_per_channel(w)                             # the tied table (row-chunked in the reference)
    ids = torch.randint(0, V, (T,), device="cuda", generator=g, dtype=torch.int32)
    ids[0] = 0; ids[-1] = V - 1
    y = _embed(ids, q, s)
    yd = _embed(ids, q, s, decode=True)
    assert torch.equal(y, yd)
    if O.ref_lib_path().exists():
        R = O.ref_lib()
        yr = torch.empty_like(y)
        rc = R.milaref_token_embedding_forward_bf16_qfp8(G.p(yr), G.p(ids), G.p(q), G.p(s), 1, T, C, ctypes.c_void_p(G.stream()))
        torch.cuda.synchronize(); assert rc == 0
        assert torch.equal(y, yr), "differs from the reference kernel"
    if V <= 5000:
        yo = O.token_embedding_qfp8(ids.cpu().numpy(), G.u8(q), G.f32(s))
        np.testing.assert_array_equal(G.bits_of(y), yo)
    # reference formula pin (TokenEmbedding.Cuda.cpp:555-577)
    rows = w[ids.long()].float(); sc = s[ids.long()][:, None]
    absmax = rows.abs().amax(dim=1, keepdim=True)
    d448 = torch.full_like(absmax, 448.0)                        # tensor divisor: a true IEEE division (x / scalar multiplies by 1/448)
    assert torch.equal(sc, torch.where(absmax > 0, torch.div(absmax, d448), torch.ones_like(absmax)))
    assert bool(((y.float() - rows).abs() <= 0.07 * rows.abs() + 0.004 * sc).all())


def test_tied_table_serves_embedding_and_lm_head():
    """One FP8 table + one scale vector: gathering row v equals projecting the one-hot... i.e. the lm_head's
    dequantised weight row v (Gemma.ixx:139-147: tied lm_head is Linear<Cuda,BF16,PerChannelFp8<>>)."""
    V, C = 2048, 512
    w = (torch.randn((V, C), device="cuda") / C ** 0.5).to(torch.bfloat16)
    q, s = quantize_fp8_per_channel(w)
    ids = torch.arange(0, V, 97, device="cuda", dtype=torch.int32)
    emb = _embed(ids, q, s)
    deq = (q.view(torch.float8_e4m3fn).float() * s[:, None])[ids.long()].to(torch.bfloat16)
    assert torch.equal(emb, deq)
    # lm_head logits of an embedded token against its own row: x . w_v with x = emb_v  (sanity of the shared scales)
    x = emb[:4].contiguous()
    logits = linear_forward(x, q, s, PerChannelFp8()).float()
    ref = x.float() @ (q.view(torch.float8_e4m3fn).float() * s[:, None]).t()
    assert H.rel_err_rowabs(logits.cpu().numpy(), ref.cpu().numpy()) <= 1e-2


def test_embedding_argument_errors():
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.milab200_token_embedding_forward_bf16_qfp8(one, one, one, one, 1, 4, 60, None) == _lib.E_BAD_SHAPE
    assert L.milab200_token_embedding_decode_bf16_qfp8(None, one, one, one, 4, 64, None) == _lib.E_INVALID_ARGUMENT
