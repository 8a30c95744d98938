#!/usr/bin/env python
"""Bring-up check of decode_mx4.cu on tiny cases (one-hot activations pick out weight columns)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from mila_b200 import _lib
from mila_b200.linear import PerGroupFp4, linear_forward, quantize_fp4_per_group
E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0]
def deq(q, s):
    lut = torch.tensor(E2M1, device=q.device)
    w = torch.stack((lut[(q & 0xF).long()], lut[(q >> 4).long()]), dim=-1).reshape(q.shape[0], -1)
    return w * s.repeat_interleave(128, dim=1)
pol = PerGroupFp4(128)
bad = 0
for (N, K, M, kind) in [(128, 128, 1, "onehot"), (128, 256, 1, "onehot"), (128, 512, 1, "onehot"), (128, 128, 1, "ones"),
                        (128, 256, 1, "randn"), (256, 1024, 3, "randn"), (3840, 4096, 4, "randn"), (3840, 15360, 1, "randn"),
                        (30720, 3840, 2, "randn")]:
    torch.manual_seed(1)
    w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    q, s = quantize_fp4_per_group(w, 128)
    if kind == "onehot":
        x = torch.zeros((M, K), device="cuda", dtype=torch.bfloat16); x[0, min(K - 1, 77)] = 1.0
    elif kind == "ones":
        x = torch.ones((M, K), device="cuda", dtype=torch.bfloat16)
    else:
        x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    try:
        y = linear_forward(x, q, s, pol).float()
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL launch", N, K, M, kind, e); bad += 1; continue
    ref = x.float() @ deq(q, s).t()
    den = torch.maximum(ref.abs(), 1e-2 * ref.abs().amax(dim=1, keepdim=True))
    err = float(((y - ref).abs() / den).max())
    ok = err <= 1e-2
    bad += (not ok)
    print(("OK  " if ok else "BAD "), N, K, M, kind, _lib.last_kernel(), "err=%.4g" % err, "y[0,:4]=", y[0, :4].tolist(), "ref=", ref[0, :4].tolist(), flush=True)
print("bad cases:", bad)
