"""GPU tests of the tensor-parallel path.  The multi-rank check (tests/tp_check.py) needs >= 2 GPUs and is
launched with torchrun; on a 1-GPU box only the world-size-1 context is exercised (the host-side logic
of N > 1 is covered on CPU by tests/test_tp_gloo.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from mila_b200.linear import PerChannelFp8, PerGroupFp4, linear_forward, quantize_fp4_per_group, quantize_fp8_per_channel
from mila_b200.tp import TpGroup

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("policy", [PerChannelFp8(), PerGroupFp4(128)], ids=["fp8", "fp4g128"])
def test_world_size_one_context_is_the_plain_forward(policy):
    tp = TpGroup(None, max_out_features=4096)
    N, K, M = 512, 1024, 3
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = quantize_fp8_per_channel(w) if isinstance(policy, PerChannelFp8) else quantize_fp4_per_group(w, 128)
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    assert torch.equal(tp.rowparallel_forward(x, q, s, policy), linear_forward(x, q, s, policy))
    tp.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_fused_allreduce_matches_single_gpu():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "tp_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "TP_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
