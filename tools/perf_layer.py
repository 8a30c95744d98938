#!/usr/bin/env python
"""Decode Linears of whole transformer layers back to back (graph replay, distinct weights per layer):
us per layer and achieved GB/s.  usage: perf_layer.py [gemma|llama8b|llama70b] [M] [layers]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200.linear import PerChannelFp8, PerGroupFp4  # noqa: E402
from mila_b200.stack import LayerChain  # noqa: E402

MODELS = {
    "gemma": ([(3840, 8192), (4096, 3840), (3840, 30720), (15360, 3840)], PerGroupFp4(128), 24),
    "llama8b": ([(4096, 6144), (4096, 4096), (4096, 28672), (14336, 4096)], PerChannelFp8(), 16),
    "llama70b": ([(8192, 10240), (8192, 8192), (8192, 57344), (28672, 8192)], PerGroupFp4(128), 6),
}
name = sys.argv[1] if len(sys.argv) > 1 else "gemma"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shapes, pol, layers = MODELS[name]
if len(sys.argv) > 3: layers = int(sys.argv[3])
ch = LayerChain(shapes, layers, pol, M)
ch.capture()
for _ in range(3): ch.step()
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.step(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
us_layer = best * 1e3 / layers
print(json.dumps({"model": name, "M": M, "layers": layers, "us_per_layer": round(us_layer, 2),
                  "GBps": round(ch.algorithmic_bytes_per_step() / layers / us_layer / 1e3, 1),
                  "weights_GB": round(ch.weight_bytes() / 1e9, 2), "launches_per_layer": ch.launches_per_step // layers}))
