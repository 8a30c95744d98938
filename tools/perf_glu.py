#!/usr/bin/env python
"""gate|up Linear + gated activation: fused epilogue vs the two-kernel sequence (Linear, then the activation
kernel), CUDA-graph replay over distinct weight copies (> 2x L2), us per (gate_up + activation).
usage: perf_glu.py [M,M,...]   (default 1,4,16; e.g. 2048 for the batched kernel's fused epilogue)"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from bench import ClockSampler  # noqa: E402  (NVML clock record on every line)
from mila_b200 import _lib  # noqa: E402
from mila_b200.linear import (GLU_GEGLU_TANH, GLU_SWIGLU, PerChannelFp8, PerGroupFp4, glu_forward,  # noqa: E402
                              linear_forward, linear_glu_forward)

CASES = [("gemma_gate_up_fp4_geglu", PerGroupFp4(128), 15360, 3840, GLU_GEGLU_TANH),
         ("llama8b_gate_up_fp8_swiglu", PerChannelFp8(), 14336, 4096, GLU_SWIGLU),
         ("llama70b_gate_up_fp4_swiglu", PerGroupFp4(128), 28672, 8192, GLU_SWIGLU)]
for name, pol, H, K, kind in CASES:
    fp8 = isinstance(pol, PerChannelFp8)
    per = 2 * H * (K if fp8 else K // 2)
    copies = max(4, int(2.5 * 126e6 / per) + 1)
    ws = []
    for _ in range(copies):
        q = torch.randint(0, 256, (2 * H, K if fp8 else K // 2), dtype=torch.uint8, device="cuda")
        if fp8: q[(q & 0x7F) == 0x7F] = 0
        s = torch.rand((2 * H,) if fp8 else (2 * H, K // 128), device="cuda") * 0.01 + 0.001
        ws.append((q, s))
    for M in ([int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else (1, 4, 16)):
        if M > 16: copies = 2
        x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
        gu = torch.empty((M, 2 * H), device="cuda", dtype=torch.bfloat16)
        out = torch.empty((M, H), device="cuda", dtype=torch.bfloat16)
        res, kern = {}, {}
        sampler = ClockSampler(torch.cuda.current_device()).start()
        for mode in ("fused", "two_step"):
            def step(i):
                q, s = ws[i % copies]
                if mode == "fused":
                    linear_glu_forward(x, q, s, pol, kind, None, out, gu)
                else:
                    linear_forward(x, q, s, pol, None, gu); glu_forward(gu, kind, out)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                step(0); step(1)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    for i in range(2 * copies): step(i)
                for _ in range(2): g.replay()
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
            res[mode] = best / (2 * copies) * 1e3
            kern[mode] = _lib.last_kernel()
        clocks = sampler.stop()
        print(json.dumps({"case": name, "M": M, "fused_us": round(res["fused"], 2), "two_step_us": round(res["two_step"], 2),
                          "speedup": round(res["two_step"] / res["fused"], 3), "kernels": kern, "clocks": clocks}), flush=True)
