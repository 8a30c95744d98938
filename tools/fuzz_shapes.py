#!/usr/bin/env python
"""Random-shape sweep of the forward entry points against a torch FP32 GEMM over exactly dequantised weights
(the dequantisation itself is pinned to the oracle by tests/test_gpu_prefill.py).  Prints failures; exit code 1
if any case exceeds the parity gate (row-abs relative error 1e-2)."""
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402
from mila_b200.linear import PerChannelFp8, PerGroupFp4, linear_forward, quantize_fp4_per_group, quantize_fp8_per_channel  # noqa: E402

E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0]


def deq(pol, q, s):
    if isinstance(pol, PerChannelFp8):
        return q.view(torch.float8_e4m3fn).float() * s[:, None]
    g = pol.kQuantizationGroupSize
    lut = torch.tensor(E2M1, device=q.device)
    w = torch.stack((lut[(q & 0xF).long()], lut[(q >> 4).long()]), dim=-1).reshape(q.shape[0], -1)
    return w * s.repeat_interleave(g, dim=1)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = random.Random(seed)
torch.backends.cuda.matmul.allow_tf32 = False
bad = 0
kernels = {}
for case in range(n):
    pol = rng.choice([PerChannelFp8(), PerGroupFp4(128), PerGroupFp4(128), PerGroupFp4(64)])
    K = rng.choice([128, 256, 384, 512, 640, 1024, 1152, 2048, 3840, 4096, 7168, 8192]) if rng.random() < 0.9 else 64 * rng.randint(1, 40)
    if isinstance(pol, PerGroupFp4) and K % pol.kQuantizationGroupSize: K = (K // 128 + 1) * 128
    N = rng.choice([1, 7, 64, 128, 129, 200, 256, 1000, 2048, 3840, 4096, 8192, 14336]) if rng.random() < 0.8 else rng.randint(1, 6000)
    M = rng.choice([1, 2, 3, 4, 5, 8, 9, 16, 17, 24, 32, 33, 64, 100, 128, 129, 256, 300])
    bias = rng.random() < 0.3
    torch.manual_seed(case + 1000 * seed)
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    x = (torch.randn((M, K), device="cuda") * rng.choice([1.0, 1e-3, 50.0])).to(torch.bfloat16)
    b = (torch.randn((N,), device="cuda") * 0.1).to(torch.bfloat16) if bias else None
    if isinstance(pol, PerChannelFp8): q, s = quantize_fp8_per_channel(w)
    else: q, s = quantize_fp4_per_group(w, pol.kQuantizationGroupSize)
    try:
        y = linear_forward(x, q, s, pol, b).float()
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL launch", type(pol).__name__, getattr(pol, "kQuantizationGroupSize", 0), N, K, M, bias, e); bad += 1; continue
    kernels[_lib.last_kernel()] = kernels.get(_lib.last_kernel(), 0) + 1
    ref = x.float() @ deq(pol, q, s).t()
    if b is not None: ref = ref + b.float()
    den = torch.maximum(ref.abs(), 1e-2 * ref.abs().amax(dim=1, keepdim=True))
    den = torch.where(den == 0, torch.ones_like(den), den)
    err = float(((y - ref).abs() / den).max())
    if not (err <= 1e-2) or not bool(torch.isfinite(y).all()):
        print("BAD", type(pol).__name__, getattr(pol, "kQuantizationGroupSize", 0), "N", N, "K", K, "M", M, "bias", bias, _lib.last_kernel(), "err", err); bad += 1
print("cases", n, "bad", bad)
for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]): print(f"  {v:4d}  {k}")
sys.exit(1 if bad else 0)
