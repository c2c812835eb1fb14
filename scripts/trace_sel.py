"""Phase timeline of the cluster select kernel (library built with `make TRACE=1`)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32736
sc = (torch.rand(1, 32, n, device=dev) * 1e-3).to(torch.bfloat16)
for _ in range(3):
    idx = _native.pool_topk(sc, 7, "avgpool", 2016)[0]
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH); buf = (ctypes.c_longlong * 16)(); raw.kvc_debug_sel_trace(buf)
t = list(buf)
names = ["start", "loaded+pooled", "sums done", "minmax done", "thresholds done", "scan done", "compacted", "sorted", "written"]
for i in range(1, 9):
    print(f"{names[i]:16s} +{t[i] - t[i-1]:7d} cycles  (t={t[i] - t[0]})")
for i, nm in ((9, "minmax: local part done"), (10, "minmax: cluster barrier passed"), (11, "round 1: counted"),
              (12, "round 1: cluster barrier passed"), (13, "compaction: stores issued")):
    print(f"{nm:32s} t={t[i] - t[0]}")
g = (ctypes.c_longlong * 8)(); raw.kvc_debug_sel_gt(g); g = list(g)
print("globaltimer (ns, relative to CTA 0's start): start of CTAs 0-3:", [x - g[0] for x in g[:4]], " local min/max done:", [x - g[0] for x in g[4:]])
tl = (ctypes.c_longlong * 16)(); raw.kvc_debug_sel_tail(tl); tl = list(tl)
print("CTA 3 last thread (cycles since its first stamp): ", [x - tl[0] for x in tl[:5]], " thread 0:", [x - tl[8] for x in tl[8:13]])
