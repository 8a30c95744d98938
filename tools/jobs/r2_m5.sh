#!/bin/bash
# session m, job 5 (diagnostic): where does the fused batched GLU epilogue lose its time? dbg 1 = no exchange, 2 = exchange + trivial math, 4 = exchange only
O=gpurun_out; mkdir -p $O
cat > /tmp/one.py <<'P'
import sys, json, torch
sys.path.insert(0, '.')
from mila_b200 import _lib
from mila_b200.linear import GLU_SWIGLU, GLU_GEGLU_TANH, PerChannelFp8, PerGroupFp4, linear_glu_forward, linear_forward
for name, pol, H, K, kind in [("llama8b_fp8_swiglu", PerChannelFp8(), 14336, 4096, GLU_SWIGLU), ("gemma_fp4_geglu", PerGroupFp4(128), 15360, 3840, GLU_GEGLU_TANH)]:
    fp8 = isinstance(pol, PerChannelFp8); M = 2048
    q = torch.randint(0, 256, (2 * H, K if fp8 else K // 2), dtype=torch.uint8, device="cuda")
    if fp8: q[(q & 0x7F) == 0x7F] = 0
    s = torch.rand((2 * H,) if fp8 else (2 * H, K // 128), device="cuda") * 0.01 + 0.001
    x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    gu = torch.empty((M, 2 * H), device="cuda", dtype=torch.bfloat16); out = torch.empty((M, H), device="cuda", dtype=torch.bfloat16)
    for mode in ("fused", "linear_only"):
        f = (lambda: linear_glu_forward(x, q, s, pol, kind, None, out, gu)) if mode == "fused" else (lambda: linear_forward(x, q, s, pol, None, gu))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        print(name, mode, round(e0.elapsed_time(e1) / 10 * 1e3, 1), "us", _lib.last_kernel(), flush=True)
P
for dbg in 0 1 2 4; do echo "== MILAB200_GLU_DBG=$dbg"; MILAB200_GLU_DBG=$dbg timeout 200 python /tmp/one.py 2>&1 | tail -4; done
