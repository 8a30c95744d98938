#!/usr/bin/env python
"""Per-shape decode microbenchmark (development tool; bench.py is the contract benchmark).

For every (format, K->N, M) it rotates over enough distinct weight copies to exceed 2x L2, captures the
launcher calls in a CUDA graph and reports us/launch, algorithmic GB/s and the fraction of the measured
HBM peak.  With --ref it also times Mila's own M=1 matvec kernels (oracle/_ref, recompiled for sm_100a)
in the same harness.  Prints one JSON line per case.
"""
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
    "fp8": [("llama8b_gate", 4096, 14336), ("llama8b_down", 14336, 4096), ("gemma_lm_head", 3840, 262144)],
    "fp4": [("gemma_qkv", 3840, 8192), ("gemma_o", 4096, 3840), ("gemma_gate_up", 3840, 30720),
            ("gemma_down", 15360, 3840), ("llama70b_up", 8192, 28672), ("llama70b_down", 28672, 8192)],
}


def p(t): return None if t is None else ctypes.c_void_p(t.data_ptr())


def time_graph(fn_list, iters=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fn_list[:2]: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fn_list: f()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best / len(fn_list) * 1e3     # us per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="fp8,fp4")
    ap.add_argument("--m", default="1,2,4,8,16")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--launches", type=int, default=48)
    args = ap.parse_args()
    L = _lib.lib()
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    R = None
    if args.ref:
        from oracle import oracle as O
        R = O.ref_lib()
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for fmt in args.fmt.split(","):
        for (name, K, N) in SHAPES[fmt]:
            if args.only and args.only not in name: continue
            wbytes = N * K if fmt == "fp8" else N * K // 2
            copies = max(2, min(args.launches, -(-400_000_000 // wbytes)))
            ws, ss = [], []
            for c in range(copies):
                q = torch.randint(0, 256, (N, K if fmt == "fp8" else K // 2), dtype=torch.uint8, device="cuda")
                if fmt == "fp8": q[(q & 0x7F) == 0x7F] = 0
                s = torch.rand((N,) if fmt == "fp8" else (N, K // 128), device="cuda") * 0.01 + 0.001
                ws.append(q); ss.append(s)
            for M in [int(v) for v in args.m.split(",")]:
                x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
                y = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
                def mk(i):
                    w, s = ws[i % copies], ss[i % copies]
                    if fmt == "fp8":
                        return lambda: _lib.check(L.milab200_w8a16_gemm(p(y), p(x), p(w), p(s), None, M, K, N, st()), "x")
                    return lambda: _lib.check(L.milab200_fp4a16_gemm(p(y), p(x), p(w), p(s), None, M, K, N, 128, st()), "x")
                n_l = max(args.launches, copies)
                sampler = ClockSampler(torch.cuda.current_device()).start()
                us = time_graph([mk(i) for i in range(n_l)], iters=12)
                clocks = sampler.stop()
                sbytes = 4 * N if fmt == "fp8" else 4 * N * K // 128
                alg = wbytes + sbytes + 2 * M * (K + N)
                out = {"fmt": fmt, "shape": name, "K": K, "N": N, "M": M, "us": round(us, 3),
                       "GBps": round(alg / us / 1e3, 1), "frac_measured_peak": round(alg / us / 1e3 / peak, 4),
                       "tok_per_s": round(M / us * 1e6), "kernel": _lib.last_kernel(), "copies": copies, "clocks": clocks}
                if R is not None and M == 1:
                    def mkr(i):
                        w, s = ws[i % copies], ss[i % copies]
                        if fmt == "fp8":
                            return lambda: R.milaref_matvec_decode_bf16_qfp8(p(y), p(x), p(w), p(s), None, K, N, st())
                        return lambda: R.milaref_matvec_decode_bf16_qfp4(p(y), p(x), p(w), p(s), None, K, N, 128, st())
                    usr = time_graph([mkr(i) for i in range(n_l)])
                    out["ref_us"] = round(usr, 3); out["ref_GBps"] = round(alg / usr / 1e3, 1)
                print(json.dumps(out), flush=True)
            del ws, ss
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
