"""Timeline of the fused scoring kernel (library built with `make TRACE=1`): per-CTA globaltimer stamps."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768
Ks = [torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
Q = torch.randn(1, 32, L, 128, device=dev).to(torch.bfloat16)
for i in range(8):
    s = _native.window_scores(Q, Ks[i % 4], 32, "sum", 3)
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH)
n = 144
buf = (ctypes.c_longlong * (n * 16))()
raw.kvc_debug_ftrace(buf, n * 16)
t = [list(buf[i * 16:(i + 1) * 16]) for i in range(n)]
t0 = min(r[0] for r in t)
names = ["start", "Qstaged", "p1done.wg0", "p1done.wg1", "p1done.wg2", "published", "barrier", "crow", "p2done.wg0",
         "p2done.wg1", "p2done.wg2", "end", "tma.p1last", "tma.p2last"]
import statistics
print("slot            min      median      max   (ns after the first CTA started)")
for k, nm in enumerate(names):
    v = [r[k] - t0 for r in t if r[k] > 0]
    if v:
        print(f"{nm:12s} {min(v):8d} {int(statistics.median(v)):10d} {max(v):8d}")
for c in (0, 1, 17, 70, 143):
    print("cta", c, [(x - t0) if x > 0 else -1 for x in t[c][:14]])
tb = (ctypes.c_uint * (4 * 64 * 8))()
raw.kvc_debug_tile_trace(tb)
for ci, cta in enumerate((0, 70)):
    c = 0 if ci == 0 else 2
    base = min(tb[(c * 64 + 0) * 8 + k] for k in range(7) if tb[(c * 64 + 0) * 8 + k] > 0)
    print(f"--- CTA {cta}: per tile (SM cycles after the first stamp): tma_issue kfull tempty mma_commit | epi_got epi_loaded epi_done")
    for j in range(30):
        r = [tb[(c * 64 + j) * 8 + k] for k in range(7)]
        print(f"  tile {j:2d}: " + " ".join(f"{((x - base) & 0xffffffff) if x > 0 else -1:7d}" for x in r[:4]) + " | " + " ".join(f"{((x - base) & 0xffffffff) if x > 0 else -1:7d}" for x in r[4:]))
