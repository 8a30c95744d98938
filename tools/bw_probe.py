#!/usr/bin/env python
"""Drive the read-only access-pattern probes (mila_b200/csrc/bw_probe.cu)."""
import ctypes
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402

L = _lib.lib()
L.milab200_probe_bw.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]
L.milab200_probe_bw.restype = ctypes.c_int
NAMES = {0: "linear", 1: "tile16x64_d4", 2: "tile16x64_d8", 3: "bulk16x256x3", 4: "bulk16x512x2",
         5: "bulk16x1024x2_4w", 6: "tile8x128_d4", 7: "row512_d8"}
out = torch.zeros(4, dtype=torch.int32, device="cuda")
for (rows, row_bytes) in [(14336, 4096), (4096, 14336), (30720, 1920), (3840, 2048), (262144, 3840)]:
    total = rows * row_bytes
    copies = max(2, min(24, -(-500_000_000 // total)))
    bufs = [torch.randint(0, 255, (total,), dtype=torch.uint8, device="cuda") for _ in range(copies)]
    for pat in sorted(NAMES):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        def launch(i):
            rc = L.milab200_probe_bw(ctypes.c_void_p(bufs[i % copies].data_ptr()), rows, row_bytes, pat,
                                          ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
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
        print(json.dumps({"rows": rows, "row_bytes": row_bytes, "MB": round(total / 1e6, 1), "pattern": NAMES[pat],
                          "us": round(us, 2), "GBps": round(total / us / 1e3, 1)}), flush=True)
    del bufs; torch.cuda.empty_cache()
