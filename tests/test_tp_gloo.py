"""CPU tests (gloo, world_size 2) of the host side of the tensor-parallel path: shard algebra against the
oracle (column shards are row slices, row shards are K slices holding whole FP4 groups and — for FP8 — the
unsharded row scales), the handle exchange, and the rank-ordered sum that the fused
all-reduce performs.  No CUDA calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_helpers as H
from mila_b200 import _lib
from mila_b200.linear import PerChannelFp8, PerGroupFp4
from mila_b200.tp import column_shard, exchange_handles, row_shard, shard_bounds
from oracle import oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, tmpdir=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. handle exchange: every rank sees every rank's 64-byte handle in rank order
        mine = bytes([rank + 1]) * 64
        hs = exchange_handles(mine)
        assert hs == [bytes([r + 1]) * 64 for r in range(world)]

        # 2. the rank-ordered FP32 sum of K-shard partials == the unsharded oracle result (to FP32 noise),
        #    and it is the same on every rank
        N, K, M = 48, 512, 3
        w = H.xavier_weights_bf16(N, K, seed=5)
        x = H.activations_bf16(M, K, seed=6)
        for policy in (PerChannelFp8(), PerGroupFp4(128)):
            if isinstance(policy, PerChannelFp8):
                qq, ss = O.quantize_fp8_per_channel(w)
            else:
                qq, ss = O.quantize_fp4_per_group(w, 128)
            qt, st = torch.from_numpy(qq), torch.from_numpy(ss)
            q_r, s_r = row_shard(qt, st, policy, world, rank)
            ks = shard_bounds(K, world, rank, 128)
            x_r = np.ascontiguousarray(x[:, ks])
            if isinstance(policy, PerChannelFp8):
                wf = O.dequant_fp8(q_r.numpy(), np.ones(N, np.float32))          # scale applied after the sum
            else:
                wf = O.dequant_fp4(q_r.numpy(), s_r.numpy(), 128)
            part = torch.from_numpy(O.bf16_bits_to_f32(x_r).astype(np.float32) @ wf.T.astype(np.float32))
            parts = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(parts, part)
            total = torch.zeros_like(part)
            for p in parts:                                                       # rank order, as the kernel adds
                total += p
            if isinstance(policy, PerChannelFp8):
                total = total * st[None, :]
                _, ref = O.linear_forward_fp8(x, qq, ss, None)
            else:
                _, ref = O.linear_forward_fp4(x, qq, ss, 128, None)
            assert H.rel_err_rowabs(total.numpy(), ref) <= 1e-4
            tots = [torch.empty_like(total) for _ in range(world)]
            dist.all_gather(tots, total)
            assert all(torch.equal(t, total) for t in tots)
        # 3. shards straight from a pre-quantized artifact: rank 0 writes the file once, every rank maps it and reads
        #    ONLY its slice (artifact.readLinearShard); the row-parallel sum over artifact shards is the unsharded result
        if tmpdir is not None:
            from mila_b200 import artifact as A
            path = os.path.join(tmpdir, "tp.safetensors")
            policy = PerGroupFp4(128)
            qq, ss = O.quantize_fp4_per_group(w, 128)
            bias_bits = O.f32_to_bf16_bits(np.linspace(-1.0, 1.0, N, dtype=np.float32))
            if rank == 0:
                wr = A.SafeTensorsWriter(path)
                wr.setMetadata(A.kMilaQuantizationMetadataKey, policy.tag)
                wr.declareTensor("down.weight", "UINT8", qq.shape); wr.declareTensor("down.weight_scale", "FP32", ss.shape)
                wr.declareTensor("down.bias", "BF16", (N,))
                wr.beginData(); wr.writeTensorData("down.weight", qq); wr.writeTensorData("down.weight_scale", ss)
                wr.writeTensorData("down.bias", bias_bits); wr.close()
            dist.barrier()
            with A.ArtifactReader(path) as r:
                q_r, s_r, b_r = A.readLinearShard(r, "down", policy, world, rank, "row")
                q_c, s_c, _ = A.readLinearShard(r, "down", policy, world, rank, "column")
            # row-parallel: EVERY rank holds the full bias (TpGroup.rowparallel_forward adds it exactly once on either
            # route: after the cross-rank sum in the fused epilogue, masked to rank 0 ahead of the NCCL sum)
            assert q_r.shape == (N, K // 2 // world) and q_c.shape == (N // world, K // 2)
            assert np.array_equal(b_r.view(torch.int16).numpy().view(np.uint16), bias_bits)
            ks = shard_bounds(K, world, rank, 128)
            part = torch.from_numpy(O.bf16_bits_to_f32(np.ascontiguousarray(x[:, ks])).astype(np.float32)
                                    @ O.dequant_fp4(q_r.numpy(), s_r.numpy(), 128).T.astype(np.float32))
            bias_f = torch.from_numpy(O.bf16_bits_to_f32(bias_bits).astype(np.float32))
            # the two routes of rowparallel_forward, restated on the host: NCCL route = bias on rank 0 before the sum,
            # fused route = bias after the sum on every rank; both give sum + bias once, identical on all ranks
            nccl_like = part + (bias_f if rank == 0 else 0.0)
            dist.all_reduce(part); dist.all_reduce(nccl_like)
            fused_like = part + bias_f
            _, ref_b = O.linear_forward_fp4(x, qq, ss, 128, bias_bits)
            assert H.rel_err_rowabs(fused_like.numpy(), ref_b) <= 1e-4 and H.rel_err_rowabs(nccl_like.numpy(), ref_b) <= 1e-4
            both = [torch.empty_like(fused_like) for _ in range(world)]
            dist.all_gather(both, fused_like)
            assert all(torch.equal(t, fused_like) for t in both)
            _, ref = O.linear_forward_fp4(x, qq, ss, 128, None)
            assert H.rel_err_rowabs(part.numpy(), ref) <= 1e-4
            # column shards: every rank's rows of the output equal the matching rows of the unsharded result
            rs = shard_bounds(N, world, rank)
            _, ref_c = O.linear_forward_fp4(x, q_c.numpy(), s_c.numpy(), 128, None)
            assert np.array_equal(ref_c, ref[:, rs])
        q.put((rank, "ok"))
    except Exception as e:                                                       # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo(tmp_path):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shards_are_slices_of_the_unsharded_quantisation():
    N, K, world = 64, 1024, 4
    w = H.xavier_weights_bf16(N, K, seed=11)
    q4, s4 = O.quantize_fp4_per_group(w, 128)
    q8, s8 = O.quantize_fp8_per_channel(w)
    for rank in range(world):
        # column-parallel: quantising the row slice alone gives the same bytes (per-row / per-group formats)
        rs = shard_bounds(N, world, rank)
        qa, sa = O.quantize_fp4_per_group(np.ascontiguousarray(w[rs]), 128)
        cq, cs = column_shard(torch.from_numpy(q4), torch.from_numpy(s4), world, rank)
        np.testing.assert_array_equal(cq.numpy(), qa); np.testing.assert_array_equal(cs.numpy(), sa)
        qb, sb = O.quantize_fp8_per_channel(np.ascontiguousarray(w[rs]))
        cq, cs = column_shard(torch.from_numpy(q8), torch.from_numpy(s8), world, rank)
        np.testing.assert_array_equal(cq.numpy(), qb); np.testing.assert_array_equal(cs.numpy(), sb)
        # row-parallel FP4: groups never straddle shards -> independent quantisation of the K slice is identical
        ks = shard_bounds(K, world, rank, 128)
        qc, sc = O.quantize_fp4_per_group(np.ascontiguousarray(w[:, ks]), 128)
        rq, rsc = row_shard(torch.from_numpy(q4), torch.from_numpy(s4), PerGroupFp4(128), world, rank)
        np.testing.assert_array_equal(rq.numpy(), qc); np.testing.assert_array_equal(rsc.numpy(), sc)
        # row-parallel FP8: bytes are a slice of the UNSHARDED quantisation, scales stay whole-row
        rq, rsc = row_shard(torch.from_numpy(q8), torch.from_numpy(s8), PerChannelFp8(), world, rank)
        np.testing.assert_array_equal(rq.numpy(), q8[:, ks]); np.testing.assert_array_equal(rsc.numpy(), s8)
    with pytest.raises(_lib.InvalidArgument):
        shard_bounds(1000, 3, 0)
    with pytest.raises(_lib.InvalidArgument):
        row_shard(torch.from_numpy(q4), torch.from_numpy(s4), PerGroupFp4(128), 16, 0)   # 64-k shards < one group


def test_rowparallel_entry_argument_errors():
    import ctypes
    L = _lib.lib()
    one = ctypes.c_void_p(16)
    assert L.milab200_w8a16_gemm_rowparallel(one, one, one, one, None, 1, 128, 128, None, None) == _lib.E_INVALID_ARGUMENT
    ctx = ctypes.c_void_p()
    assert L.milab200_tp_create(3, 2, 128, ctypes.byref(ctx)) == _lib.E_INVALID_ARGUMENT
    assert L.milab200_tp_create(0, 9, 128, ctypes.byref(ctx)) == _lib.E_INVALID_ARGUMENT
    assert L.milab200_tp_handle_bytes() == 64
