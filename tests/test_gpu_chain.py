"""GPU parity of the chained decode kernel (csrc/decode_chain.cu, milab200_chain_*): a list of dependent decode Linears
as ONE persistent launch must give what the per-Linear launchers give on the same tensors — bit for bit wherever the
FP32 summation order is the same (whole-k row tiles of any height), within one BF16 rounding of the FP32
dequantise-then-GEMM result everywhere (gate 1e-2) — and the same bits on every replay."""
import numpy as np
import pytest
import torch

import gpu_util as G
import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import (GLU_GEGLU_TANH, GLU_SWIGLU, PerChannelFp8, PerGroupFp4, linear_forward, linear_glu_forward,
                              quantize_fp4_per_group, quantize_fp8_per_channel)
from mila_b200.stack import DecodeChain, LinearStack
from mila_b200.tp import dequantize_fp32, rel_err_rowabs
from oracle import oracle as O

pytestmark = pytest.mark.gpu
POLICIES = [PerChannelFp8(), PerGroupFp4(128)]
IDS = ["fp8", "fp4g128"]


def _quant(policy, N, K, seed):
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    w = (torch.randn((N, K), device="cuda", generator=gen) / K ** 0.5).to(torch.bfloat16)
    return quantize_fp8_per_channel(w) if isinstance(policy, PerChannelFp8) else quantize_fp4_per_group(w, 128)


def _bias(N, seed):
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    return (torch.randn((N,), device="cuda", generator=gen) * 0.25).to(torch.bfloat16)


def _fp32_ref(x, q, s, policy, bias=None):
    y = x.float() @ dequantize_fp32(q, s, policy).t()
    return y if bias is None else y + bias.float()


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
@pytest.mark.parametrize("M", [1, 3, 8, 11, 16])
def test_chain_of_three_linears_matches_per_linear_launches(policy, M):
    """A: 512 -> 1024 (+bias), B: 1024 -> 384, C: 384 -> 640 (+bias); every entry reads the previous entry's output."""
    dims = [(512, 1024, True), (1024, 384, False), (384, 640, True)]
    ws = [(_quant(policy, N, K, 10 + i), _bias(N, 20 + i) if b else None) for i, (K, N, b) in enumerate(dims)]
    gen = torch.Generator(device="cuda"); gen.manual_seed(M)
    x = torch.randn((M, 512), device="cuda", generator=gen).to(torch.bfloat16)
    outs = [torch.zeros((M, N), device="cuda", dtype=torch.bfloat16) for (_, N, _) in dims]
    entries, src = [], x
    for ((q, s), b), o in zip(ws, outs):
        entries.append({"x": src, "weight": q, "scales": s, "bias": b, "out": o})
        src = o
    chain = DecodeChain(entries, policy, M, "cuda:0")
    chain.forward(); torch.cuda.synchronize()
    assert _lib.last_kernel().startswith("decode_chain_kernel"), _lib.last_kernel()
    got = [o.clone() for o in outs]
    # per-Linear launches on the chain's own intermediate activations: each entry is checked on identical inputs
    src = x
    for i, (((q, s), b), g_) in enumerate(zip(ws, got)):
        want = linear_forward(src, q, s, policy, b)
        torch.cuda.synchronize()
        ref = _fp32_ref(src, q, s, policy, b)
        assert rel_err_rowabs(g_.float(), ref) <= 1e-2, (i, rel_err_rowabs(g_.float(), ref))
        assert rel_err_rowabs(g_.float(), want.float()) <= 8e-3, i          # at most one BF16 ulp apart (k-split order)
        src = g_
    # replays give the same bits; a changed input changes the output
    for o in outs: o.zero_()
    chain.forward(); torch.cuda.synchronize()
    for o, g_ in zip(outs, got):
        assert torch.equal(o, g_)
    chain.close()


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
@pytest.mark.parametrize("kind", [GLU_GEGLU_TANH, GLU_SWIGLU], ids=["geglu", "swiglu"])
@pytest.mark.parametrize("M", [1, 5, 16])
def test_chain_gate_up_glu_then_down_equals_fused_launches_bit_for_bit(policy, kind, M):
    """The MLP dataflow (fc_gate_up + GLU -> fc_down, Gemma.Block.ixx:347-349): whole-k row tiles add in the same order
    whatever their height, so the chained GLU entry equals the stand-alone fused launcher bit for bit; ragged H."""
    K, Hh = 768, 14400                                    # enough rows for the stand-alone fused launcher; ragged last tile
    qgu, sgu = _quant(policy, 2 * Hh, K, 1)
    gen = torch.Generator(device="cuda"); gen.manual_seed(7 * M)
    x = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    bgu = _bias(2 * Hh, 3)
    h = torch.zeros((M, Hh), device="cuda", dtype=torch.bfloat16)
    chain = DecodeChain([{"x": x, "weight": qgu, "scales": sgu, "bias": bgu, "out": h, "glu": kind}], policy, M, "cuda:0")
    chain.forward(); torch.cuda.synchronize()
    # same operand scheme on both sides: FP4 at M <= 2 is the packed-nibble kind::mxf4 scheme in the chain and in the
    # stand-alone launcher alike; everything else the E4M3-plane kind::f8f6f4 scheme
    mx = isinstance(policy, PerGroupFp4) and M <= 2
    want = linear_glu_forward(x, qgu, sgu, policy, kind, bgu)
    torch.cuda.synchronize()
    same_scheme = _lib.last_kernel().startswith("decode_mx4_kernel" if mx else "decode_tc_kernel")
    if same_scheme:
        assert torch.equal(h, want), _lib.last_kernel()
    else:
        # the stand-alone launcher picked the other operand scheme for this shape: both are exact products with FP32
        # accumulation, so the BF16 results agree to one ulp
        assert rel_err_rowabs(h.float(), want.float()) <= 8e-3, _lib.last_kernel()
    chain.close()


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
def test_chain_balanced_decomposition_and_split_k_pairs(policy):
    """Llama-8B shapes: gate/up are cut into one wave of equal-height tiles without a k split, down (few rows, long k)
    into two k halves per tile — the two CTAs of a cluster meeting over distributed shared memory — and the result still
    meets the parity gate against the FP32 reference and the per-Linear launch."""
    M, hidden, ffn = 4, 4096, 14336
    (qg, sg), (qd, sd) = _quant(policy, ffn, hidden, 5), _quant(policy, hidden, ffn, 6)
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    x = torch.randn((M, hidden), device="cuda", generator=gen).to(torch.bfloat16)
    g = torch.zeros((M, ffn), device="cuda", dtype=torch.bfloat16)
    y = torch.zeros((M, hidden), device="cuda", dtype=torch.bfloat16)
    bd = _bias(hidden, 9)
    chain = DecodeChain([{"x": x, "weight": qg, "scales": sg, "out": g},
                         {"x": g, "weight": qd, "scales": sd, "bias": bd, "out": y}], policy, M, "cuda:0")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    d0, d1 = chain.describe(0), chain.describe(1)
    assert d0["ksplits"] == 1 and sms - 2 <= d0["tiles"] <= sms, d0
    assert d1["ksplits"] == 2 and sms - 2 <= 2 * d1["tiles"] <= sms, d1
    chain.forward(); torch.cuda.synchronize()
    g_want = linear_forward(x, qg, sg, policy)
    torch.cuda.synchronize()
    assert _lib.last_kernel().startswith("decode_tc_kernel"), _lib.last_kernel()          # (M = 4: E4M3-plane scheme)
    assert torch.equal(g, g_want)                          # whole-k tiles: same FP32 order as the 128-row launch
    assert rel_err_rowabs(y.float(), _fp32_ref(g, qd, sd, policy, bd)) <= 1e-2
    assert rel_err_rowabs(y.float(), linear_forward(g, qd, sd, policy, bd).float()) <= 8e-3
    y1 = y.clone(); y.zero_()
    chain.forward(); torch.cuda.synchronize()
    assert torch.equal(y, y1)
    chain.close()


@pytest.mark.parametrize("policy", POLICIES, ids=IDS)
@pytest.mark.parametrize("M", [1, 16])
def test_stack_in_chain_mode_equals_launch_mode_under_graph_replay(policy, M):
    """LinearStack (the bench harness): 3 layers of (gate, up, down), launch mode vs chain mode on the same seeds, captured
    and replayed three times with fresh inputs: identical to launch mode wherever the order is the same, and within
    the parity gate of the FP32 reference computed from the chain's own activations."""
    hidden, ffn, layers = 1024, 2816, 3
    a = LinearStack(hidden, ffn, layers, policy, M, "cuda:0", mode="launches")
    b = LinearStack(hidden, ffn, layers, policy, M, "cuda:0", mode="chain")
    a.capture(); b.capture()
    assert b.launches_per_step == 1 and a.launches_per_step >= 3 * layers
    gen = torch.Generator(device="cuda"); gen.manual_seed(11)
    for it in range(3):
        x = torch.randn((M, hidden), device="cuda", generator=gen).to(torch.bfloat16)
        a.set_input(x); b.set_input(x)
        ya = a.step().clone(); yb = b.step().clone()
        torch.cuda.synchronize()
        # three BF16-rounded layers: a one-ulp difference in a split-k layer propagates, so compare per layer below
        assert torch.isfinite(yb.float()).all()
        gate, up, down = b.w[-1]
        ref = _fp32_ref(b.g, down.weight, down.scales, policy)           # last layer on the chain's own activations
        assert rel_err_rowabs(yb.float(), ref) <= 1e-2
        # launch mode on the same inputs: three BF16-rounded (FP4: coarsely quantized) layers amplify a one-ulp difference of
        # a split-k layer, so across modes only the direction of the result is comparable
        cos = torch.nn.functional.cosine_similarity(yb.float().flatten(), ya.float().flatten(), dim=0)
        assert float(cos) > 0.999, (it, float(cos))
    b.chain.close()


def test_chain_argument_errors():
    import ctypes
    L = _lib.lib()
    ctx = ctypes.c_void_p()
    arr = (_lib.ChainLinear * 1)()
    assert L.milab200_chain_create(arr, 0, 1, ctypes.byref(ctx)) == _lib.E_INVALID_ARGUMENT
    assert L.milab200_chain_create(arr, 1, 17, ctypes.byref(ctx)) == _lib.E_BAD_SHAPE            # batched: per-Linear entries
    assert L.milab200_chain_create(arr, 1, 1, ctypes.byref(ctx)) == _lib.E_INVALID_ARGUMENT      # null tensors
    q, s = _quant(PerGroupFp4(128), 256, 256, 1)
    x = torch.zeros((1, 256), device="cuda", dtype=torch.bfloat16); y = torch.zeros((1, 256), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.MilaB200Error):
        DecodeChain([{"x": x, "weight": q, "scales": s, "out": y}], PerGroupFp4(32), 1, "cuda:0")   # unsupported group
    with pytest.raises(_lib.InvalidArgument):
        DecodeChain([{"x": x, "weight": q, "scales": s, "out": y, "depends_on": 0}], PerGroupFp4(128), 1, "cuda:0")
    assert L.milab200_chain_forward(None, None) == _lib.E_INVALID_ARGUMENT


@pytest.mark.parametrize("M", [1, 2])
def test_chain_four_way_k_split_on_the_packed_nibble_scheme(M):
    """Gemma-12B FP4 shapes at M <= 2: the down projection (few rows, long k) is cut into FOUR k quarters per tile — two
    cluster pairs meeting over distributed shared memory, the second pair handing its sum to the first as tagged FP32 words
    through L2 — so that a unit is 104 rows high instead of 52 (the block-scaled MMAs cost the same whatever the height).
    Parity against the FP32 reference and the per-Linear launch, hand-off to a following entry, same bits on every run."""
    policy = PerGroupFp4(128)
    hidden, ffn = 3840, 15360
    (qg, sg), (qd, sd), (qg2, sg2) = _quant(policy, ffn, hidden, 15), _quant(policy, hidden, ffn, 16), _quant(policy, ffn, hidden, 17)
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    x = torch.randn((M, hidden), device="cuda", generator=gen).to(torch.bfloat16)
    g = torch.zeros((M, ffn), device="cuda", dtype=torch.bfloat16)
    y = torch.zeros((M, hidden), device="cuda", dtype=torch.bfloat16)
    g2 = torch.zeros((M, ffn), device="cuda", dtype=torch.bfloat16)
    bd = _bias(hidden, 19)
    chain = DecodeChain([{"x": x, "weight": qg, "scales": sg, "out": g},
                         {"x": g, "weight": qd, "scales": sd, "bias": bd, "out": y},
                         {"x": y, "weight": qg2, "scales": sg2, "out": g2}], policy, M, "cuda:0")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    d1 = chain.describe(1)
    assert d1["ksplits"] == 4 and sms - 4 <= 4 * d1["tiles"] <= sms, d1
    chain.forward(); torch.cuda.synchronize()
    assert _lib.last_kernel() == "decode_chain_kernel<fp4g128,packed,t2>"
    y1, g21 = y.clone(), g2.clone()
    assert rel_err_rowabs(y.float(), _fp32_ref(g, qd, sd, policy, bd)) <= 1e-2
    assert rel_err_rowabs(g2.float(), _fp32_ref(y, qg2, sg2, policy)) <= 1e-2
    want = linear_forward(g, qd, sd, policy, bd); torch.cuda.synchronize()
    assert rel_err_rowabs(y.float(), want.float()) <= 8e-3
    for _ in range(3):
        y.zero_(); g2.zero_()
        chain.forward(); torch.cuda.synchronize()
        assert torch.equal(y, y1) and torch.equal(g2, g21)
    chain.close()


@pytest.mark.parametrize("M", [1, 2, 4])
@pytest.mark.parametrize("name,policy,hidden,shard", [
    ("llama70b_fp4_world8", PerGroupFp4(128), 8192, 28672 // 8),       # down: k = 3584 = 3.5 packed-nibble units of 1024 k
    ("llama70b_fp4_world4", PerGroupFp4(128), 8192, 28672 // 4),
    ("llama8b_fp8_world8", PerChannelFp8(), 4096, 14336 // 8),
    ("llama8b_fp8_world4", PerChannelFp8(), 4096, 14336 // 4),
    ("gemma12b_fp4_world8", PerGroupFp4(128), 3840, 15360 // 8),       # k = 1920 = 15 groups
])
def test_chain_on_tensor_parallel_shard_shapes(name, policy, hidden, shard, M):
    """The per-rank Linears of the tensor-parallel stacks at world 4 and 8 (gate / up column shards hidden -> ffn / world, the
    down projection's k shard ffn / world -> hidden), chained on ONE GPU: the tile / unit / k-split decomposition of those
    shapes (short k that is not a whole number of packed-nibble units, few rows) is what a rank of the 8-GPU bench runs; the
    exchange itself is checked by tests/tp_check.py under torchrun."""
    (qg, sg), (qu, su), (qd, sd), (qg2, sg2) = (_quant(policy, shard, hidden, 31), _quant(policy, shard, hidden, 32),
                                                _quant(policy, hidden, shard, 33), _quant(policy, shard, hidden, 34))
    gen = torch.Generator(device="cuda"); gen.manual_seed(7 + M)
    x = torch.randn((M, hidden), device="cuda", generator=gen).to(torch.bfloat16)
    g, u, g2 = (torch.zeros((M, shard), device="cuda", dtype=torch.bfloat16) for _ in range(3))
    y = torch.zeros((M, hidden), device="cuda", dtype=torch.bfloat16)
    chain = DecodeChain([{"x": x, "weight": qg, "scales": sg, "out": g},
                         {"x": x, "weight": qu, "scales": su, "out": u, "depends_on": -1},
                         {"x": g, "weight": qd, "scales": sd, "out": y, "depends_on": 0},
                         {"x": y, "weight": qg2, "scales": sg2, "out": g2}], policy, M, "cuda:0")
    chain.forward(); torch.cuda.synchronize()
    assert _lib.last_kernel().startswith("decode_chain_kernel"), _lib.last_kernel()
    got = [t.clone() for t in (g, u, y, g2)]
    for out, src, q, s in ((g, x, qg, sg), (u, x, qu, su), (y, g, qd, sd), (g2, y, qg2, sg2)):
        assert rel_err_rowabs(out.float(), _fp32_ref(src, q, s, policy)) <= 1e-2, name
        want = linear_forward(src, q, s, policy); torch.cuda.synchronize()
        assert rel_err_rowabs(out.float(), want.float()) <= 8e-3, name
    for _ in range(2):
        for t in (g, u, y, g2): t.zero_()
        chain.forward(); torch.cuda.synchronize()
        for t, w in zip((g, u, y, g2), got):
            assert torch.equal(t, w)
    chain.close()


@pytest.mark.parametrize("mode", ["launches", "chain"])
def test_stack_host_buffer_pass_equals_the_device_resident_pass(mode):
    """LinearStack.forward_host (what bench.py times as `e2e`): pinned H2D of the activations, the stack, D2H of the result,
    replayed as one graph — the same bits as set_input + step on the same activations, for the stack's own staging buffer and
    for a caller's host tensor."""
    policy, M, hidden, ffn = PerChannelFp8(), 2, 1024, 2816
    st = LinearStack(hidden, ffn, 2, policy, M, "cuda:0", mode=mode)
    st.capture()
    gen = torch.Generator(device="cpu"); gen.manual_seed(3)
    for it in range(3):
        xh = torch.randn((M, hidden), generator=gen).to(torch.bfloat16)
        if it == 0:
            st.x_host.copy_(xh); yh = st.forward_host()
        else:
            yh = st.forward_host(xh)
        torch.cuda.current_stream().synchronize()
        got = yh.clone()
        st.set_input(xh.to("cuda"))
        want = st.step().clone(); torch.cuda.synchronize()
        assert torch.equal(got, want.cpu()), it
        assert float(got.float().abs().max()) > 0
    if st.chain is not None:
        st.chain.close()
