#!/usr/bin/env python
"""Per-shape batched-forward (prefill) microbenchmark (development tool).

For every (format, K->N, M): CUDA-graph of `launches` forwards rotating over 2 weight copies, CUDA
events, best of `iters`.  Reports us per forward (act_split + GEMM), useful TFLOP/s = 2 M N K / t and
the fraction of the measured BF16 dense peak (MEASURED_PEAKS.json bf16_tflops; the two-plane E4M3 MMA
does 2x the FP8-rate work per useful flop, so BF16 dense is the matching denominator)."""
import argparse
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from mila_b200 import _lib  # noqa: E402
from bench import ClockSampler  # noqa: E402  (NVML clock record on every line)

SHAPES = {
    "fp8": [("llama8b_gate", 4096, 14336), ("llama8b_down", 14336, 4096)],
    "fp4": [("gemma_qkv", 3840, 8192), ("gemma_o", 4096, 3840), ("gemma_gate_up", 3840, 30720),
            ("gemma_down", 15360, 3840), ("llama70b_up", 8192, 28672)],
}


def p(t): return None if t is None else ctypes.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="fp8,fp4")
    ap.add_argument("--m", default="2048")
    ap.add_argument("--only", default="")
    ap.add_argument("--launches", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    L = _lib.lib()
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text())["bf16_tflops"] if peaks.exists() else 1590.0
    for fmt in args.fmt.split(","):
        for name, K, N in SHAPES[fmt]:
            if args.only and args.only not in name:
                continue
            ws = []
            for _ in range(2):
                q = torch.randint(0, 256, (N, K if fmt == "fp8" else K // 2), dtype=torch.uint8, device="cuda")
                if fmt == "fp8": q[(q & 0x7F) == 0x7F] = 0
                s = torch.rand((N,) if fmt == "fp8" else (N, K // 128), device="cuda") * 0.01 + 0.001
                ws.append((q, s))
            for M in [int(v) for v in args.m.split(",")]:
                x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
                y = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
                _lib.check(L.milab200_reserve_prefill(M, K), "reserve")

                def fwd(i, st):
                    q, s = ws[i % 2]
                    if fmt == "fp8":
                        rc = L.milab200_w8a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, st)
                    else:
                        rc = L.milab200_fp4a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, 128, st)
                    _lib.check(rc, "fwd")

                strm = torch.cuda.Stream()
                with torch.cuda.stream(strm):
                    st = ctypes.c_void_p(strm.cuda_stream)
                    fwd(0, st); fwd(1, st)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=strm):
                        for i in range(args.launches): fwd(i, st)
                    for _ in range(2): g.replay()
                    torch.cuda.synchronize()
                    best = 1e9
                    sampler = ClockSampler(torch.cuda.current_device()).start()
                    for _ in range(args.iters):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                        best = min(best, e0.elapsed_time(e1))
                    clocks = sampler.stop()
                us = best / args.launches * 1e3
                tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
                print(json.dumps({"fmt": fmt, "shape": name, "K": K, "N": N, "M": M, "us": round(us, 2),
                                  "TFLOPs": round(tf, 1), "frac_bf16_peak": round(tf / peak, 4),
                                  "kernel": _lib.last_kernel(),
                                  "clocks": clocks}), flush=True)


if __name__ == "__main__":
    main()
