#!/usr/bin/env python
"""Run the chained decode kernel a few times without CUDA graphs (the command ncu wraps).
usage: ncu_chain_case.py [layers] [tokens] [workload]   — `layers` x (gate, up, down) of the workload (bench.py WORKLOADS,
default llama3.1-8b-mlp-fp8) in ONE launch"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402
from bench import WORKLOADS  # noqa: E402
from mila_b200.linear import PerChannelFp8, PerGroupFp4  # noqa: E402
from mila_b200.stack import LinearStack  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = sys.argv[3] if len(sys.argv) > 3 else "llama3.1-8b-mlp-fp8"
hidden, ffn, _, pol = WORKLOADS[wl]
st = LinearStack(hidden, ffn, layers, PerChannelFp8() if pol == "fp8" else PerGroupFp4(128), M, "cuda:0", mode="chain")
st.set_input(torch.randn((M, hidden), device="cuda").to(torch.bfloat16))
torch.cuda.synchronize()
for _ in range(4):
    st.step()
torch.cuda.synchronize()
print("ok", _lib.last_kernel(), "algorithmic bytes per launch", st.algorithmic_bytes_per_step(), float(st.h[0].float().abs().mean()))
