#!/usr/bin/env python
"""Is the unpacked-E2M1 (16U4_ALIGN16B) TMA path bound by bytes in flight or by the TMA unit itself?
Same probe as tools/tma_probe.py (libmila_b200_probes.so), a few decisive variants: u8 vs u4 boxes at equal shared-memory
ring size and at equal HBM bytes in flight."""
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

L = ctypes.CDLL(str(ROOT / "mila_b200" / "libmila_b200_probes.so"))
c_i, c_p, c_l = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64
L.milab200_probe_tma.argtypes = [c_p, c_l, c_l, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
L.milab200_probe_tma.restype = c_i
# (u4, R, C, S)
VARIANTS = [(0, 128, 1, 12), (0, 128, 1, 6), (0, 128, 1, 3), (1, 128, 1, 12), (1, 128, 1, 6), (1, 128, 4, 3), (0, 128, 4, 3),
            (0, 64, 1, 24), (1, 64, 1, 24), (0, 256, 1, 6), (1, 256, 1, 6)]
SHAPES = [(262144, 3840), (14336, 4096), (28672, 4096)]
for (rows, row_bytes) in SHAPES:
    total = rows * row_bytes
    copies = max(2, min(24, -(-500_000_000 // total)))
    bufs = [torch.randint(0, 255, (total,), dtype=torch.uint8, device="cuda") for _ in range(copies)]
    for (u4, R, C, S) in VARIANTS:
        def launch(i):
            rc = L.milab200_probe_tma(c_p(bufs[i % copies].data_ptr()), rows, row_bytes, u4, R, C, S, 0, 1, 3, 148, 0, None, None,
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
        print(json.dumps({"rows": rows, "row_bytes": row_bytes, "MB": round(total / 1e6, 1), "u4": u4, "R": R, "C": C, "S": S,
                          "smem_ring_KB": R * C * S * 128 // 1024, "hbm_in_flight_KB": R * C * S * (64 if u4 else 128) // 1024,
                          "us": round(us, 2), "GBps": round(total / us / 1e3, 1)}), flush=True)
    del bufs; torch.cuda.empty_cache()
