#!/usr/bin/env python
"""Per-unit role timeline of CTA 0 of the pre-split kind::mxf4 decode kernel (bring-up tool).
usage: mx8_timeline.py K N M   — clock64 stamps relative to CTA 0's entry, one row per unit:
P = TMA producer, M = MMA issuer, E = epilogue thread 0."""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402

K, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
L = _lib.lib()
if not hasattr(L, "milab200_diag_set_tc_prof"):
    raise SystemExit("this tool needs the diagnostics build: make -C mila_b200/csrc diag && MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_diag.so python " + sys.argv[0])
L.milab200_diag_set_tc_prof.argtypes = [ctypes.c_void_p]
L.milab200_diag_set_tc_prof.restype = None
p = lambda t: ctypes.c_void_p(t.data_ptr())
ws = []
for _ in range(8):
    q = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device="cuda")
    s = torch.rand((N, K // 128), device="cuda") * 0.01 + 0.001
    ws.append((q, s))
x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
y = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
prof = torch.zeros(64 * 16 + 148 * 4 + 148 + 64, dtype=torch.int64, device="cuda")


def go(i):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    q, s = ws[i % 8]
    rc = L.milab200_fp4a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, 128, st)
    assert rc == 0, rc


for i in range(3): go(i)
torch.cuda.synchronize()
L.milab200_diag_set_tc_prof(p(prof))
g = torch.cuda.CUDAGraph()
s_ = torch.cuda.Stream()
with torch.cuda.stream(s_):
    go(3)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for i in range(6): go(i)
    g.replay()
torch.cuda.synchronize()
L.milab200_diag_set_tc_prof(None)
print(_lib.last_kernel())
cta = prof.cpu()[1024:1024 + 148 * 4].view(148, 4)
cta = cta[cta[:, 0] > 0]
print("graph of 6 (last kernel's stamps): CTAs %d, kernel span %d ns, entry spread %d ns, exit spread %d ns"
      % (cta.shape[0], int(cta[:, 3].max() - cta[:, 0].min()), int(cta[:, 0].max() - cta[:, 0].min()),
         int(cta[:, 3].max() - cta[:, 3].min())))
t = prof.cpu()[:1024].view(64, 16).tolist()
names = ["P:empty", "P:A", "P:B", "-", "-", "-", "M:tempty", "M:full", "M:commit", "E:start", "E:tfull", "E:ld", "E:done", "M:mmas", "-"]
print("unit " + " ".join(n.rjust(9) for n in names))
for i, row in enumerate(t):
    if row[1] == 0 and i > 0: break
    print(f"{i:4d} " + " ".join(str(v).rjust(9) for v in row[:15]))
