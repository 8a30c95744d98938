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
        if M > 16 or force_nccl:
            # bias must be added once: rank 0 carries it
            linear_forward(x, weight, scales, policy, bias if self.rank == 0 else None, out)
            dist.all_reduce(out, group=self.group)
            return out
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(x.device):
            if isinstance(policy, PerChannelFp8):
                rc = L.milab200_w8a16_gemm_rowparallel(p(out), p(x), p(weight), p(scales), p(bias), M, K, N, self._ctx, st)
            else:
                rc = L.milab200_fp4a16_gemm_rowparallel(p(out), p(x), p(weight), p(scales), p(bias), M, K, N,
                                                        policy.kQuantizationGroupSize, self._ctx, st)
        _lib.check(rc, "rowparallel_forward")
        return out

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
