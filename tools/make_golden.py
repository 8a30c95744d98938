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

# gated activations behind the gate|up Linear (Activations/{Geglu,Swiglu}/Kernels, compiled unmodified):
# a sweep of BF16 gate values x a few up values, plus seeded random rows
import ctypes  # noqa: E402
R = O.ref_lib()
sweep = np.concatenate([np.linspace(-12, 12, 2049, dtype=np.float32), np.array([0.0, -0.0, 1e-30, -1e-30, 50, -50, 1e4, -1e4, 88.0, -88.0, 100.0, -100.0], np.float32)])
sweep = sweep[: (len(sweep) // 8) * 8]
Hh = len(sweep)
ups = np.array([1.0, -0.75, 3.0, 1e-3], np.float32)
xg = np.zeros((len(ups) + 2, 2 * Hh), np.float32)
for i, u in enumerate(ups):
    xg[i, :Hh] = sweep; xg[i, Hh:] = u
rng = np.random.default_rng(2024)
xg[len(ups):, :] = rng.standard_normal((2, 2 * Hh)).astype(np.float32) * 2.5
xb = O.f32_to_bf16_bits(xg)
xd = G.bf16_tensor(xb, "cuda")
rec = {"x_bits": xb}
for kind, fn in ((1, R.milaref_geglu_forward_bf16), (2, R.milaref_swiglu_forward_bf16)):
    y = torch.empty((xb.shape[0], Hh), dtype=torch.bfloat16, device="cuda")
    rc = fn(G.p(y), G.p(xd), y.numel(), Hh, ctypes.c_void_p(G.stream()))
    torch.cuda.synchronize(); assert rc == 0
    rec["geglu_y_bits" if kind == 1 else "swiglu_y_bits"] = G.bits_of(y)
np.savez_compressed(out / "glu_sweep.npz", **rec)
print("glu_sweep", {k: v.shape for k, v in rec.items()})
# RMSNorm (Normalizations/RmsNorm/Kernels/RmsNorm.Bf16.cu, compiled unmodified): plain and Gemma (1 + w) variants
rec = {}
for tag, (M, K, off, with_b, eps) in {"plain": (6, 768, 0.0, True, 1e-5), "gemma": (4, 3840, 1.0, False, 1e-6)}.items():
    x = H.activations_bf16(M, K, seed=41 + M) ; x = O.f32_to_bf16_bits(O.bf16_bits_to_f32(x) * np.float32(2.5))
    w = O.f32_to_bf16_bits((np.random.default_rng(5).standard_normal(K) * 0.3).astype(np.float32))
    b = O.f32_to_bf16_bits((np.random.default_rng(6).standard_normal(K) * 0.1).astype(np.float32)) if with_b else None
    xd, wd = G.bf16_tensor(x, "cuda"), G.bf16_tensor(w, "cuda")
    bd = G.bf16_tensor(b, "cuda") if b is not None else None
    y = torch.empty_like(xd)
    rc = R.milaref_rmsnorm_forward_bf16(G.p(y), None, G.p(xd), G.p(wd), G.p(bd), M, 1, K, ctypes.c_float(eps), ctypes.c_float(off),
                                        ctypes.c_void_p(G.stream()))
    torch.cuda.synchronize(); assert rc == 0
    rec[f"{tag}_x"], rec[f"{tag}_w"], rec[f"{tag}_y"] = x, w, G.bits_of(y)
    if b is not None: rec[f"{tag}_b"] = b
    rec[f"{tag}_eps"], rec[f"{tag}_off"] = np.float32(eps), np.float32(off)
np.savez_compressed(out / "rmsnorm_ref.npz", **rec)
print("rmsnorm_ref", {k: getattr(v, "shape", v) for k, v in rec.items()})

# PerGroupInt4 forward (cuda_w4a16_gemm, CudaW4A16Gemm.cu:88-197, compiled unmodified): symmetric and asymmetric
rec = {}
for tag, (N, K, M, g, asym) in {"sym_g128": (40, 512, 3, 128, False), "asym_g64": (24, 256, 5, 64, True)}.items():
    rng = np.random.default_rng(len(tag))
    w = rng.integers(0, 256, (N, K // 2), dtype=np.uint8)
    sc = (rng.random((N, K // g), dtype=np.float32) * 0.02 + 0.001).astype(np.float32)
    z = rng.integers(0, 256, (N, K // g // 2), dtype=np.uint8) if asym else None
    x = H.activations_bf16(M, K, seed=17)
    bias = O.f32_to_bf16_bits(np.linspace(-0.5, 0.5, N, dtype=np.float32))
    y = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    zd = torch.from_numpy(z).cuda() if z is not None else None
    xd, wd, sd, bd = G.bf16_tensor(x, "cuda"), torch.from_numpy(w).cuda(), torch.from_numpy(sc).cuda(), G.bf16_tensor(bias, "cuda")
    rc = R.milaref_w4a16_gemm(G.p(y), G.p(xd), G.p(wd), G.p(sd), G.p(zd), G.p(bd), M, K, N, g, ctypes.c_void_p(G.stream()))
    torch.cuda.synchronize(); assert rc == 0
    rec[f"{tag}_w"], rec[f"{tag}_s"], rec[f"{tag}_x"], rec[f"{tag}_bias"], rec[f"{tag}_y"] = w, sc, x, bias, G.bits_of(y)
    if z is not None: rec[f"{tag}_z"] = z
    rec[f"{tag}_g"] = np.int32(g)
np.savez_compressed(out / "int4_ref.npz", **rec)
print("int4_ref", {k: getattr(v, "shape", v) for k, v in rec.items()})
print("device:", torch.cuda.get_device_name(0))
