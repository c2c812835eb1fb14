"""Per-CTA timeline of the dynamically scheduled scoring kernel (needs `make TRACE=1 -C kvcache_factory_b200/csrc`
and KVC_LIB_PATH=kvcache_factory_b200/libkvc_b200_trace.so).  usage: trace_dyn.py [L] [Hq] [Hkv]"""
import os, sys, ctypes, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
Hq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Hkv = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda:0")
Ks = [torch.randn(1, Hkv, L, 128, device=dev).to(torch.bfloat16) for _ in range(6)]
Q = torch.randn(1, Hq, 32, 128, device=dev).to(torch.bfloat16)
for i in range(8):
    s = _native.window_scores(Q, Ks[i % 6], 32, "sum", _native.IMPL_TCGEN05_DYN, seq_len=L)
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH)
n = 148
buf = (ctypes.c_longlong * (n * 8))()
raw.kvc_debug_dyn_trace(buf, n * 8)
t = [list(buf[i * 8:(i + 1) * 8]) for i in range(n)]
t0 = min(r[0] for r in t if r[0] > 0)
names = ["Q staged", "first tile", "chunk 1 published", "chunk 2 published", "x folded", "y folded", "end"]
print("stamp            min   median      max   (ns after the first CTA started)")
for k, nm in enumerate(names):
    v = [r[k] - t0 for r in t if r[k] > 0]
    if v:
        print(f"{nm:14s} {min(v):7d} {int(statistics.median(v)):8d} {max(v):8d}")
n1 = [r[7] // 1000 for r in t]; n2 = [r[7] % 1000 for r in t]
print("pass-1 tiles per CTA: min %d max %d sum %d;  pass-2: min %d max %d sum %d" % (min(n1), max(n1), sum(n1), min(n2), max(n2), sum(n2)))
for c in (0, 1, 8, 70, 147):
    print("cta", c, [(x - t0) if x > 1000 else x for x in t[c]])

if hasattr(raw, "kvc_debug_dyn_warp_trace"):
    wb = (ctypes.c_uint * (4 * 64 * 4))()
    raw.kvc_debug_dyn_warp_trace(wb)
    base = min(wb[(g * 64 + 0) * 4 + 0] for g in range(4))
    print("--- CTA 0, the four epilogue warps of one SM sub-partition (one per warpgroup): per tile, SM cycles: seen loaded done  (x4 warpgroups)")
    for j in range(34):
        row = []
        for g in range(4):
            r = [wb[(g * 64 + j) * 4 + k] for k in range(4)]
            row.append(" ".join(f"{(x - base) & 0xffffffff:6d}" for x in r[:3]))
        code = wb[(0 * 64 + j) * 4 + 3]
        kind = "stop" if code == 0xffffffff else ("P2" if code & (1 << 27) else "P1") + (".y" if code & (1 << 26) else ".x")
        print(f"  {j:2d} {kind:5s} | " + " | ".join(row))
if not hasattr(raw, "kvc_debug_dyn_tile_trace"):
    sys.exit(0)          # the per-tile stamps exist only in experimental builds of scores_dyn.cu
tb = (ctypes.c_uint * (2 * 64 * 8))()
raw.kvc_debug_dyn_tile_trace(tb)
for ci, cta in enumerate((0, 70)):
    base = tb[(ci * 64 + 0) * 8 + 0]
    print(f"--- CTA {cta}: per tile, SM cycles after the first TMA issue: tma_issue | mma kfull_seen, tempty_seen | epi(warp 2) tfull_seen loaded done | code")
    prev = 0
    for j in range(40):
        r = [tb[(ci * 64 + j) * 8 + k] for k in range(8)]
        code = r[6]
        kind = "stop" if code == 0xffffffff else ("P2" if code & (1 << 27) else "P1") + (".y" if code & (1 << 26) else ".x") + (" L" if code & (1 << 28) else "")
        d = [((x - base) & 0xffffffff) for x in r[:6]]
        print(f"  {j:2d}: {d[0]:7d} | {d[1]:7d} {d[2]:7d} | {d[3]:7d} {d[4]:7d} {d[5]:7d}  (+{d[5] - prev:5d}) wg3 done {(r[7] - base) & 0xffffffff:7d} {kind} tile {code & ((1 << 26) - 1)}")
        prev = d[5]
