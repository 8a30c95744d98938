#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own CUDA kernels (oracle/_ref, compiled
unmodified from /root/reference by oracle/Makefile) on deterministic inputs.  Runs on a GPU box:

    gpurun -- 'python tools/make_golden.py gpurun_out/golden'      # then copy gpurun_out/golden/* to tests/golden/

The vectors pin the CPU oracle (oracle/mila_oracle.c) to the reference bit for bit: packed FP8 bytes,
packed E2M1 nibbles, FP32 scales (tests/test_golden.py, runs without a GPU), and give the reference
M=1 matvec outputs the decode parity is anchored on.  Inputs are regenerated from seeds by
tests/parity_helpers.py, so only outputs (and the adversarial input bits) are stored.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import gpu_util as G  # noqa: E402
import parity_helpers as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
out.mkdir(parents=True, exist_ok=True)

CASES = {
    # name: (weights bf16 bits [N,K])
    "xavier_48x512_seed1234": H.xavier_weights_bf16(48, 512, seed=1234),
    "xavier_20x1024_seed7": H.xavier_weights_bf16(20, 1024, seed=7),
    "adversarial_16x256": H.adversarial_weights_bf16(16, 256, group=128),
    "adversarial_12x128_g64": H.adversarial_weights_bf16(12, 128, group=64),
    "reference_fixture_32x256": O.ref_weight_blob(32, 256),        # Linear.Cuda.cpp:70-75 generator
}
for name, w in CASES.items():
    N, K = w.shape
    rec = {"w_bits": w}
    q8, s8 = G.ref_quantize_fp8(w)
    rec["fp8_q"], rec["fp8_s"] = q8, s8
    for g in (128, 64):
        if K % g: continue
        q4, s4 = G.ref_quantize_fp4(w, g)
        rec[f"fp4g{g}_q"], rec[f"fp4g{g}_s"] = q4, s4
    # reference M=1 matvecs on seeded activations (finite cases only: NaN weights make outputs NaN)
    x = H.activations_bf16(1, K, seed=99)
    xd = G.bf16_tensor(x, "cuda")
    rec["x_bits"] = x
    y8 = G.ref_matvec(xd, torch.from_numpy(q8).cuda(), torch.from_numpy(s8).cuda(), 0)
    rec["fp8_y_bits"] = G.bits_of(y8)
    for g in (128, 64):
        if K % g: continue
        y4 = G.ref_matvec(xd, torch.from_numpy(rec[f"fp4g{g}_q"]).cuda(), torch.from_numpy(rec[f"fp4g{g}_s"]).cuda(), g)
        rec[f"fp4g{g}_y_bits"] = G.bits_of(y4)
    np.savez_compressed(out / f"{name}.npz", **rec)
    print(name, {k: v.shape for k, v in rec.items()})
print("device:", torch.cuda.get_device_name(0))
