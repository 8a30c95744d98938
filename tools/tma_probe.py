#!/usr/bin/env python
"""Drive the TMA tile access-pattern probes (mila_b200/csrc/tma_probe.cu)."""
import ctypes
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402

L = _lib.lib()
c_i, c_p, c_l = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64
L.milab200_probe_tma.argtypes = [c_p, c_l, c_l, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
L.milab200_probe_tma.restype = c_i

# (u4, R, C, S, mode, L, promo[, hs, tm_in_global])
VARIANTS2 = [
    (0, 128, 1, 10, 0, 1, 3, 0, 0), (0, 128, 1, 10, 0, 1, 3, 1, 0), (0, 128, 1, 10, 0, 1, 3, 2, 0), (0, 128, 1, 10, 0, 1, 3, 0, 1),
    (0, 128, 1, 10, 0, 1, 3, 1, 1),
    (0, 64, 1, 24, 0, 1, 3, 0, 0), (0, 64, 1, 24, 0, 1, 3, 1, 0), (0, 64, 1, 24, 0, 1, 3, 2, 0), (0, 64, 1, 24, 0, 1, 3, 0, 1),
    (1, 128, 1, 12, 0, 1, 3, 0, 0), (1, 128, 1, 12, 0, 1, 3, 1, 0), (1, 128, 1, 12, 0, 1, 3, 2, 0), (1, 128, 1, 12, 0, 1, 3, 0, 1),
    (0, 128, 2, 6, 0, 1, 3, 1, 0), (1, 128, 4, 3, 0, 1, 3, 1, 0),
]
VARIANTS = [
    (0, 128, 1, 10, 0, 1, 3), (0, 128, 1, 10, 0, 1, 2), (0, 128, 1, 10, 0, 1, 0),
    (0, 128, 1, 12, 1, 1, 3), (0, 128, 1, 12, 1, 4, 3), (0, 128, 1, 12, 1, 8, 3),
    (0, 128, 2, 6, 0, 1, 3), (0, 128, 4, 3, 0, 1, 3), (0, 128, 4, 3, 1, 1, 3),
    (0, 64, 4, 6, 0, 1, 3), (0, 64, 8, 3, 0, 1, 3), (0, 32, 8, 6, 0, 1, 3), (0, 32, 16, 3, 0, 1, 3),
    (0, 16, 32, 3, 0, 1, 3), (0, 256, 1, 6, 0, 1, 3), (0, 64, 1, 24, 0, 1, 3), (0, 64, 1, 24, 1, 1, 3),
    (1, 128, 1, 12, 0, 1, 3), (1, 128, 1, 12, 1, 1, 3), (1, 128, 2, 6, 0, 1, 3), (1, 128, 2, 6, 1, 1, 3),
    (1, 128, 4, 3, 0, 1, 3), (1, 64, 4, 6, 0, 1, 3), (1, 64, 4, 6, 1, 1, 3), (1, 32, 8, 6, 0, 1, 3), (1, 32, 8, 6, 1, 1, 3),
]
SHAPES = [(14336, 4096), (4096, 14336), (30720, 1920), (8192, 1920), (262144, 3840)]
if len(sys.argv) > 1 and sys.argv[1] == "hs":
    VARIANTS = VARIANTS2
    SHAPES = [(14336, 4096), (30720, 1920)]
tmbuf = torch.zeros(256, dtype=torch.uint8, device="cuda")
prof = torch.zeros(8 + 240, dtype=torch.int64, device="cuda")
PROF = len(sys.argv) > 1 and sys.argv[1] == "prof"
if PROF:
    VARIANTS = [(0, 128, 1, 10, 0, 1, 3), (0, 64, 1, 24, 0, 1, 3), (0, 256, 1, 6, 0, 1, 3), (0, 128, 2, 6, 0, 1, 3), (0, 32, 1, 32, 0, 1, 3),
                (1, 128, 1, 12, 0, 1, 3), (1, 128, 1, 24, 0, 1, 3), (1, 256, 1, 6, 0, 1, 3), (1, 256, 1, 12, 0, 1, 3), (1, 128, 2, 6, 0, 1, 3), (1, 128, 4, 6, 0, 1, 3),
                (1, 64, 1, 24, 0, 1, 3), (0, 128, 1, 10, 0, 1, 0), (1, 128, 1, 24, 0, 1, 0)]
    SHAPES = [(14336, 4096), (262144, 3840)]
for (rows, row_bytes) in SHAPES:
    total = rows * row_bytes
    copies = max(2, min(24, -(-500_000_000 // total)))
    bufs = [torch.randint(0, 255, (total,), dtype=torch.uint8, device="cuda") for _ in range(copies)]
    for var in VARIANTS:
        (u4, R, C, S, mode, Lc, promo) = var[:7]
        hs, tmg = (var[7], var[8]) if len(var) > 7 else (0, 0)
        if row_bytes % ((64 if u4 else 128) * C) != 0:
            continue
        def launch(i):
            rc = L.milab200_probe_tma(c_p(bufs[i % copies].data_ptr()), rows, row_bytes, u4, R, C, S, mode, Lc, promo,
                                           148, hs, c_p(tmbuf.data_ptr()) if tmg else None,
                                           c_p(prof.data_ptr()) if PROF else None,
                                           c_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, rc
        launch(0); torch.cuda.synchronize()
        n = max(24, copies)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n): launch(i)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        us = best / n * 1e3
        if PROF:
            pr = prof.cpu().tolist(); c = max(pr[5], 1)
            n_ = min(int(pr[5]), 120)
            d = [pr[8 + i] for i in range(n_)]
            gaps = [d[i + 1] - d[i] for i in range(n_ // 2, n_ - 1)]
            print("  producer per-stage: wait %.0f issue %.0f | first done@ %d, steady gap %.0f cycles/stage (n=%d)"
                  % (pr[0] / c, pr[2] / c, d[0] if d else -1, sum(gaps) / max(len(gaps), 1), n_))
        print(json.dumps({"rows": rows, "row_bytes": row_bytes, "MB": round(total / 1e6, 1), "u4": u4, "R": R, "C": C, "S": S,
                          "mode": mode, "L": Lc, "promo": promo, "hs": hs, "tmg": tmg, "us": round(us, 2), "GBps": round(total / us / 1e3, 1)}), flush=True)
    del bufs; torch.cuda.empty_cache()
