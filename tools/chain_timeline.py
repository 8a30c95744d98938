#!/usr/bin/env python
"""Per-layer role timeline of the chained decode kernel (bring-up tool).
usage: chain_timeline.py [workload] [tokens] [layers] [--fuse]   -> per-layer medians over CTAs, in us relative to the
layer's first producer issue: dependency resolved, first / last MMA, last-unit epilogue, signal; and the period."""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from bench import WORKLOADS  # noqa: E402
from mila_b200 import _lib  # noqa: E402
from mila_b200.linear import PerChannelFp8, PerGroupFp4  # noqa: E402
from mila_b200.stack import LinearStack  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "llama3.1-8b-mlp-fp8"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fuse = "--fuse" in sys.argv
hidden, ffn, _, pol = WORKLOADS[wl]
policy = PerChannelFp8() if pol == "fp8" else PerGroupFp4(128)
st = LinearStack(hidden, ffn, layers, policy, M, "cuda:0", mode="chain", fuse_gate_up=fuse)
st.set_input(torch.randn((M, hidden), device="cuda").to(torch.bfloat16))
n = st.chain.count
grid = ctypes.c_int()
L = _lib.lib()
prof = torch.zeros((n, 8, 148), dtype=torch.int64, device="cuda")
for _ in range(3): st.step()
torch.cuda.synchronize()
_lib.check(L.milab200_chain_set_timeline(st.chain._ctx, ctypes.c_void_p(prof.data_ptr()), ctypes.byref(grid)), "tl")
G = grid.value
prof = torch.zeros((n, 8, G), dtype=torch.int64, device="cuda")
_lib.check(L.milab200_chain_set_timeline(st.chain._ctx, ctypes.c_void_p(prof.data_ptr()), ctypes.byref(grid)), "tl")
st.step(); torch.cuda.synchronize()
_lib.check(L.milab200_chain_set_timeline(st.chain._ctx, None, None), "tl")
p = prof.cpu().double() / 1e3          # us
t0 = p[p > 0].min()
names = ["prod_first", "prod_last", "dep_ready", "mma_first", "mma_last", "signal", "epi_done"]
print(f"{wl} M={M} layers={layers} fuse={fuse} grid={G}")
print("layer  R   P  tiles |" + "".join(f"{k:>11s}" for k in names) + " | all-signalled  next-dep-ready(max)")
prev_all = None
for l in range(n):
    d = st.chain.describe(l)
    row = []
    for s in range(7):
        v = p[l, s]; v = v[v > 0]
        row.append((v.median() - t0).item() if v.numel() else float("nan"))
    sig = p[l, 5]; sig = sig[sig > 0]
    all_sig = (sig.max() - t0).item() if sig.numel() else float("nan")
    dep = p[l, 2]; dep = dep[dep > 0]
    extra = ""
    if d["ksplits"] == 2:
        ev, od = p[l, 7, 0::2], p[l, 7, 1::2]
        le = p[l, 2, 0::2]
        extra = (f"  | split: leader at exchange {(le[le > 0].median() - t0).item():.2f}, partner parked {(od[od > 0].median() - t0).item():.2f}"
                 f" (max {(od[od > 0].max() - t0).item():.2f}), leader has it {(ev[ev > 0].median() - t0).item():.2f}"
                 f" | mma_last leader {(p[l, 4, 0::2].median() - t0).item():.2f} partner {(p[l, 4, 1::2].median() - t0).item():.2f}")
    print(f"{l:4d} {d['tile_rows']:4d} {d['ksplits']:2d} {d['tiles']:5d} |" + "".join(f"{v:11.2f}" for v in row) +
          f" | {all_sig:10.2f}  {((dep.max() - t0).item() if dep.numel() else float('nan')):10.2f}" + extra)

# per-CTA lateness of the check-in relative to the layer median: systematic (same CTAs / SMs every layer) or random?
import numpy as np  # noqa: E402
sig = p[:, 5, :].numpy()                                  # [layer, cta]
ok = (sig > 0).all(axis=1)
late = sig[ok] - np.median(sig[ok], axis=1, keepdims=True)
mean_late = late.mean(axis=0); std_late = late.std(axis=0)
order = np.argsort(-mean_late)
print("per-CTA check-in lateness vs layer median (us): mean over layers, std over layers")
print("  slowest:", [(int(c), round(float(mean_late[c]), 2), round(float(std_late[c]), 2)) for c in order[:12]])
print("  fastest:", [(int(c), round(float(mean_late[c]), 2), round(float(std_late[c]), 2)) for c in order[-6:]])
print("  spread of the per-CTA means: p50 %.2f p90 %.2f max %.2f ; typical within-CTA std %.2f" %
      (np.percentile(mean_late, 50), np.percentile(mean_late, 90), mean_late.max(), np.median(std_late)))
mf = p[:, 3, :].numpy(); ml = p[:, 4, :].numpy()
span = (ml - mf)[ok]
print("  MMA span per layer (first->last issue), median over CTAs:", [round(float(v), 2) for v in np.median(span, axis=1)][:9])
print("  MMA span per CTA (mean over layers) p10 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(span.mean(axis=0), [10, 50, 90, 100])))
