#!/usr/bin/env python
"""Diagnostic (torchrun): chained column->row-parallel layers, per-layer comparison with the single-GPU chain."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from mila_b200.linear import PerChannelFp8, PerGroupFp4, linear_forward, quantize_fp4_per_group, quantize_fp8_per_channel
from mila_b200.tp import TpGroup, column_shard, row_shard

def rel(y, ref):
    ra = ref.abs().amax(dim=-1, keepdim=True); den = torch.maximum(ref.abs(), 1e-2 * ra)
    return float(((y - ref).abs() / den).max()), int((y != ref).sum())

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tp = TpGroup(dist.group.WORLD, 16384, dev)
policy = PerChannelFp8(); hidden, ffn, M = 4096, 14336, 4
g = torch.Generator(device=dev); g.manual_seed(4242)
wu = (torch.randn((ffn, hidden), device=dev, generator=g) / hidden ** 0.5).to(torch.bfloat16)
wd = (torch.randn((hidden, ffn), device=dev, generator=g) / ffn ** 0.5).to(torch.bfloat16)
qu, su = quantize_fp8_per_channel(wu); qd, sd = quantize_fp8_per_channel(wd)
qu_r, su_r = column_shard(qu, su, world, rank); qd_r, sd_r = row_shard(qd, sd, policy, world, rank)
xs = torch.randn((M, hidden), device=dev, generator=g).to(torch.bfloat16)
for mode in ("sync_each", "nosync", "fresh_out"):
    cur = xs; ref = xs
    out = torch.empty((M, hidden), device=dev, dtype=torch.bfloat16)
    h_r = torch.empty((M, ffn // world), device=dev, dtype=torch.bfloat16)
    res = []
    for l in range(3):
        if mode == "fresh_out":
            h = linear_forward(cur, qu_r, su_r, policy)
            o = tp.rowparallel_forward(h, qd_r, sd_r, policy)
        else:
            linear_forward(cur, qu_r, su_r, policy, None, h_r)
            o = tp.rowparallel_forward(h_r, qd_r, sd_r, policy, None, out)
        if mode == "sync_each":
            torch.cuda.synchronize()
        hf = linear_forward(ref, qu, su, policy)
        ref = linear_forward(hf, qd, sd, policy)
        res.append((o.clone(), ref.clone(), (h_r if mode != "fresh_out" else h).clone(), hf.clone()))
        cur = o
    torch.cuda.synchronize()
    for l, (o, rf, hh, hf) in enumerate(res):
        sl = slice(rank * (ffn // world), (rank + 1) * (ffn // world))
        print(f"rank{rank} {mode} layer{l}: out {rel(o.float(), rf.float())}  h_shard_vs_full_slice {rel(hh.float(), hf[:, sl].float())}", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
