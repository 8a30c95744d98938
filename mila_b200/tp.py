"""Tensor-parallel plumbing for the quantized Linear path.

The reference is single-GPU; this is the B200 part of the north star: QKV / gate / up column-parallel,
o_proj / down row-parallel, one process per GPU.  torch.distributed is used for ONE thing here: the
all-gather of the 64-byte CUDA-IPC handles of the per-rank exchange buffers.  The decode-time all-reduce
itself runs inside the row-parallel GEMV's epilogue over NVLink peer memory (csrc/tp.cu, decode_tc.cu);
for batched M > 16 the BF16 partials go through NCCL (bandwidth-bound regime).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .linear import PerChannelFp8, PerGroupFp4, linear_forward


def shard_bounds(total: int, world: int, rank: int, multiple: int = 1) -> slice:
    """Contiguous equal shard [rank*total/world, (rank+1)*total/world); `multiple` is the unit a shard must
    hold whole (the FP4 group size for a K split)."""
    if total % world != 0 or (total // world) % multiple != 0:
        raise _lib.InvalidArgument(f"cannot split {total} over {world} ranks in units of {multiple}")
    n = total // world
    return slice(rank * n, (rank + 1) * n)


def column_shard(weight: torch.Tensor, scales: torch.Tensor, world: int, rank: int):
    """Rows [r*N/p, (r+1)*N/p) of a quantized weight: FP4 and FP8 quantisation are per row / per (row, group),
    so the shard is a bit-exact slice and could equally be quantised per shard."""
    rs = shard_bounds(weight.shape[0], world, rank)
    return weight[rs].contiguous(), scales[rs].contiguous()


def row_shard(weight: torch.Tensor, scales: torch.Tensor, policy, world: int, rank: int):
    """K slice of a quantized weight.  FP4: whole groups, scales sliced with them (bit-exact).  FP8: the row
    scale is the absmax of the WHOLE row, so the shard is cut from the unsharded quantisation and every
    rank keeps the full scale vector."""
    if isinstance(policy, PerChannelFp8):
        ks = shard_bounds(weight.shape[1], world, rank, 16)
        return weight[:, ks].contiguous(), scales
    g = policy.kQuantizationGroupSize
    ks = shard_bounds(weight.shape[1] * 2, world, rank, g)
    return (weight[:, ks.start // 2: ks.stop // 2].contiguous(),
            scales[:, ks.start // g: ks.stop // g].contiguous())


class TpGroup:
    """One rank's view of a tensor-parallel group: the exchange buffer of the fused row-parallel all-reduce."""

    def __init__(self, group=None, max_out_features: int = 16384, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L = _lib.lib()
        ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.milab200_tp_create(self.rank, self.world, max_out_features, ctypes.byref(ctx)), "tp_create")
            self._ctx = ctx
            if self.world > 1:
                nb = L.milab200_tp_handle_bytes()
                mine = (ctypes.c_uint8 * nb)()
                _lib.check(L.milab200_tp_export(ctx, mine), "tp_export")
                handles = exchange_handles(bytes(mine), group)
                blob = (ctypes.c_uint8 * (nb * self.world)).from_buffer_copy(b"".join(handles))
                _lib.check(L.milab200_tp_connect(ctx, blob), "tp_connect")
                dist.barrier(group)                       # nobody starts pushing before everyone has mapped

    def rowparallel_forward(self, x, weight, scales, policy, bias=None, out=None, force_nccl: bool = False):
        """Row-parallel Linear on this rank's K shard; returns the full sum on every rank."""
        K = x.shape[-1]; M = x.numel() // K; N = weight.shape[0]
        if out is None:
            out = torch.empty((*x.shape[:-1], N), dtype=torch.bfloat16, device=x.device)
        if self.world == 1:
            return linear_forward(x, weight, scales, policy, bias, out)
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        if M > 16 or force_nccl:
            comm = self._nccl_comm()
            if comm:
                # the C-ABI entry a Mila caller uses: shard GEMM (bias on communicator rank 0 only) + ncclAllReduce on this stream
                with torch.cuda.device(x.device):
                    if isinstance(policy, PerChannelFp8):
                        rc = L.milab200_w8a16_gemm_rowparallel_nccl(p(out), p(x), p(weight), p(scales), p(bias), M, K, N,
                                                                    ctypes.c_void_p(comm), st)
                    else:
                        rc = L.milab200_fp4a16_gemm_rowparallel_nccl(p(out), p(x), p(weight), p(scales), p(bias), M, K, N,
                                                                     policy.kQuantizationGroupSize, ctypes.c_void_p(comm), st)
                _lib.check(rc, "rowparallel_forward (nccl)")
                return out
            # no raw communicator handle from this torch build: same arithmetic through torch.distributed
            linear_forward(x, weight, scales, policy, bias if self.rank == 0 else None, out)
            dist.all_reduce(out, group=self.group)
            return out
        with torch.cuda.device(x.device):
            if isinstance(policy, PerChannelFp8):
                rc = L.milab200_w8a16_gemm_rowparallel(p(out), p(x), p(weight), p(scales), p(bias), M, K, N, self._ctx, st)
            else:
                rc = L.milab200_fp4a16_gemm_rowparallel(p(out), p(x), p(weight), p(scales), p(bias), M, K, N,
                                                        policy.kQuantizationGroupSize, self._ctx, st)
        _lib.check(rc, "rowparallel_forward")
        return out

    def _nccl_comm(self) -> int:
        """ncclComm_t of this group as an integer (0 when the torch build does not expose it)."""
        if getattr(self, "_comm", None) is None:
            self._comm = 0
            try:
                g = self.group if self.group is not None else dist.group.WORLD
                self._comm = int(g._get_backend(self.device)._comm_ptr())
            except Exception:
                self._comm = 0
        return self._comm

    def close(self):
        if getattr(self, "_ctx", None):
            _lib.lib().milab200_tp_destroy(self._ctx)
            self._ctx = None


def exchange_handles(mine: bytes, group=None) -> list[bytes]:
    """All-gather of fixed-size opaque handles (works on any backend: gloo on CPU, nccl on GPU)."""
    world = dist.get_world_size(group)
    out: list = [None] * world
    dist.all_gather_object(out, mine, group=group)
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == len(mine) for h in out)
    return [bytes(h) for h in out]


# ---- self-check of a live tensor-parallel group against the single-GPU result --------------------

_E2M1_MAGNITUDES = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0)


def dequantize_fp32(weight: torch.Tensor, scales: torch.Tensor, policy) -> torch.Tensor:
    """FP32 image of a quantized weight, in torch (checker arithmetic for the parity record below — the formulas of
    Linear.Cuda.cpp:70-80 / CudaMatVecBias.Bf16.cu:249,461-494: E4M3 byte * row scale; E2M1 nibble, low nibble = even
    column, * group scale).  Not on any forward path."""
    if isinstance(policy, PerChannelFp8):
        return weight.view(torch.float8_e4m3fn).float() * scales.float()[:, None]
    g = policy.kQuantizationGroupSize
    lut = torch.tensor(_E2M1_MAGNITUDES, dtype=torch.float32, device=weight.device)
    lo, hi = (weight & 0xF).long(), (weight >> 4).long()

    def dec(n):
        return torch.where((n & 8) != 0, -lut[n & 7], lut[n & 7])
    w = torch.stack((dec(lo), dec(hi)), dim=-1).reshape(weight.shape[0], -1)
    return (w.reshape(w.shape[0], -1, g) * scales.float()[:, :, None]).reshape(w.shape[0], -1)


def rel_err_rowabs(y: torch.Tensor, ref: torch.Tensor) -> float:
    """The parity gate of SURVEY.md §8d: max |y - ref| / max(|ref|, 1e-2 * row absmax)."""
    y, ref = y.double(), ref.double()
    den = torch.maximum(ref.abs(), 1e-2 * ref.abs().amax(dim=-1, keepdim=True))
    den = torch.where(den == 0, torch.ones_like(den), den)
    return float(((y - ref).abs() / den).max())


def tp_parity_record(tp: TpGroup, policy, hidden: int, ffn: int, M: int, seed: int = 4242, bias: bool = True) -> dict:
    """One column-parallel -> row-parallel Linear pair on this group against the UNSHARDED layers on one GPU.
    Every rank builds the same unsharded quantisation, keeps its shard, and checks:
      * column-parallel: the shard's output rows equal the same rows of the unsharded Linear (row-abs relative error);
      * row-parallel (fed the bit-identical slice of the single-GPU activations, so the comparison isolates the
        all-reduce): the fused NVLink all-reduce result against (a) the single-GPU Linear's BF16 output and (b) the
        FP32 dequantise-then-GEMM result of the unsharded layer — the north-star parity bar, 1e-2;
      * identical bits on every rank.
    Collective: every rank must call it.  Returns the record (same on every rank)."""
    dev, world, rank = tp.device, tp.world, tp.rank
    from .stack import make_quant_weight
    up = make_quant_weight(ffn, hidden, policy, dev, seed)
    down = make_quant_weight(hidden, ffn, policy, dev, seed + 1)
    gen = torch.Generator(device=dev); gen.manual_seed(seed + 2)
    x = torch.randn((M, hidden), device=dev, generator=gen).to(torch.bfloat16)
    b = (torch.randn((hidden,), device=dev, generator=gen) * 0.1).to(torch.bfloat16) if bias else None
    up_q, up_s = column_shard(up.weight, up.scales, world, rank)
    dn_q, dn_s = row_shard(down.weight, down.scales, policy, world, rank)
    shard = shard_bounds(ffn, world, rank)

    h_full = linear_forward(x, up.weight, up.scales, policy)
    y_full = linear_forward(h_full, down.weight, down.scales, policy, b)
    h_col = linear_forward(x, up_q, up_s, policy)
    h_r = h_full[:, shard].contiguous()
    y_tp = tp.rowparallel_forward(h_r, dn_q, dn_s, policy, b).clone()
    kernel = _lib.last_kernel()
    torch.cuda.synchronize(dev)
    ref32 = h_full.float() @ dequantize_fp32(down.weight, down.scales, policy).t()
    if b is not None:
        ref32 = ref32 + b.float()
    identical = True
    if world > 1:
        gathered = [torch.empty_like(y_tp) for _ in range(world)]
        dist.all_gather(gathered, y_tp, group=tp.group)
        identical = all(torch.equal(o, y_tp) for o in gathered)
    rec = torch.tensor([rel_err_rowabs(h_col.float(), h_full[:, shard].float()),
                        rel_err_rowabs(y_tp.float(), y_full.float()),
                        rel_err_rowabs(y_tp.float(), ref32),
                        rel_err_rowabs(y_full.float(), ref32),
                        0.0 if identical else 1.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rec, op=dist.ReduceOp.MAX, group=tp.group)
    col, vs_single, vs_fp32, single_vs_fp32, differ = (float(v) for v in rec.tolist())
    return {"world": world, "policy": policy.tag, "shape": f"{hidden}->{ffn}->{hidden}", "M": M,
            "column_max_rel_err_rowabs": col, "max_rel_err_rowabs": vs_single,
            "max_rel_err_rowabs_vs_fp32_dequant_gemm": vs_fp32, "single_gpu_vs_fp32_dequant_gemm": single_vs_fp32,
            "identical_bits_across_ranks": differ == 0.0, "kernel": kernel}
