#!/usr/bin/env python
"""Launch one batched-forward case a few times without CUDA graphs (the command ncu wraps).
usage: ncu_prefill_case.py fmt K N M [launches] [glu]   (glu: N is the gate|up width 2 H, SwiGLU in the epilogue, FP8 only)"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402

fmt, K, N, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 3
glu = len(sys.argv) > 6 and sys.argv[6] == "glu"
L = _lib.lib()
p = lambda t: ctypes.c_void_p(t.data_ptr())
q = torch.randint(0, 256, (N, K if fmt == "fp8" else K // 2), dtype=torch.uint8, device="cuda")
if fmt == "fp8": q[(q & 0x7F) == 0x7F] = 0
s = torch.rand((N,) if fmt == "fp8" else (N, K // 128), device="cuda") * 0.01 + 0.001
x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
y = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
for i in range(launches):
    if glu:
        rc = L.milab200_w8a16_gemm_glu(p(y), None, p(x), p(q), p(s), None, M, K, N // 2, 2, st)
    elif fmt == "fp8":
        rc = L.milab200_w8a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, st)
    else:
        rc = L.milab200_fp4a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, 128, st)
    assert rc == 0, rc
torch.cuda.synchronize()
print("ok", _lib.last_kernel(), float(y.float().abs().mean()))
