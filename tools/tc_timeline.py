#!/usr/bin/env python
"""Per-unit role timeline of CTA 0 of the tcgen05 decode kernel (bring-up tool).
usage: tc_timeline.py fmt K N M"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from mila_b200 import _lib  # noqa: E402

fmt, K, N, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
L = _lib.lib()
if not hasattr(L, "milab200_diag_set_tc_prof"):
    raise SystemExit("this tool needs the diagnostics build: make -C mila_b200/csrc diag && MILAB200_LIB=$PWD/mila_b200/libmila_b200_linear_diag.so python " + sys.argv[0])
L.milab200_diag_set_tc_prof.argtypes = [ctypes.c_void_p]
L.milab200_diag_set_tc_prof.restype = None
p = lambda t: ctypes.c_void_p(t.data_ptr())
ws = []
for _ in range(4):
    q = torch.randint(0, 256, (N, K if fmt == "fp8" else K // 2), dtype=torch.uint8, device="cuda")
    if fmt == "fp8": q[(q & 0x7F) == 0x7F] = 0
    s = torch.rand((N,) if fmt == "fp8" else (N, K // 128), device="cuda") * 0.01 + 0.001
    ws.append((q, s))
x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
y = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
prof = torch.zeros(64 * 16 + 148 * 4 + 148 + 64, dtype=torch.int64, device="cuda")


def go(i):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    q, s = ws[i % 4]
    if fmt == "fp8":
        rc = L.milab200_w8a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, st)
    else:
        rc = L.milab200_fp4a16_gemm(p(y), p(x), p(q), p(s), None, M, K, N, 128, st)
    assert rc == 0, rc


for i in range(3): go(i)
torch.cuda.synchronize()
# two back-to-back launches inside one graph (as the decode loop runs them): the first records nothing,
# the second records; per-CTA globaltimer stamps show launch gap / prologue / tail
L.milab200_diag_set_tc_prof(p(prof))
go(3)
torch.cuda.synchronize()
cta = prof.cpu()[1024:1024 + 148 * 4].view(148, 4)
cta = cta[cta[:, 0] > 0]
t0 = int(cta[:, 0].min())
import statistics as _s
def col(j): return [int(v) - t0 for v in cta[:, j].tolist()]
print("per-CTA globaltimer (ns from first CTA entry): entry min/med/max %d/%d/%d | loop start %d/%d/%d | loop end %d/%d/%d | exit %d/%d/%d"
      % (min(col(0)), _s.median(col(0)), max(col(0)), min(col(1)), _s.median(col(1)), max(col(1)),
         min(col(2)), _s.median(col(2)), max(col(2)), min(col(3)), _s.median(col(3)), max(col(3))))
smid = prof.cpu()[1024 + 148 * 4:1024 + 148 * 4 + 148].tolist()
nb = cta.shape[0]
dur = [(int(cta[b, 2]) - int(cta[b, 1])) for b in range(nb)]
print("CTAs:", nb, " loop duration ns min/med/max:", min(dur), sorted(dur)[nb // 2], max(dur))
# PDL pair: launch two kernels back to back with the profile on both; the buffer keeps the second
prof.zero_()
g = torch.cuda.CUDAGraph()
s_ = torch.cuda.Stream()
with torch.cuda.stream(s_):
    with torch.cuda.graph(g):
        for i in range(6): go(i)
    g.replay()
torch.cuda.synchronize()
cta2 = prof.cpu()[1024:1024 + 148 * 4].view(148, 4)
cta2 = cta2[cta2[:, 0] > 0]
print("graph of 6 (last kernel's stamps): kernel span %d ns, entry spread %d ns, exit spread %d ns"
      % (int(cta2[:, 3].max() - cta2[:, 0].min()), int(cta2[:, 0].max() - cta2[:, 0].min()), int(cta2[:, 3].max() - cta2[:, 3].min())))
L.milab200_diag_set_tc_prof(None)
t = prof.cpu()[:1024].view(64, 16).tolist()
names = ["P:empty", "P:issued", "C0:start", "C1:start", "C0:arrive", "C1:arrive", "-", "M:full", "M:commit",
         "E:start", "E:tfull", "E:ld", "-", "M:mmas", "-"]
print(fmt, K, N, M, _lib.last_kernel())
print("unit " + " ".join(n.rjust(9) for n in names))
for i, row in enumerate(t):
    if row[1] == 0 and i > 0: break
    print(f"{i:4d} " + " ".join(str(v).rjust(9) for v in row[:15]))
