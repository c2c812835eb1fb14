"""Phase timeline of the cluster select kernel (library built with `make TRACE=1`)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0")
sc = (torch.rand(1, 32, 32736, device=dev) * 1e-3).to(torch.bfloat16)
for _ in range(3):
    idx = _native.pool_topk(sc, 7, "avgpool", 2016)[0]
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH); buf = (ctypes.c_longlong * 16)(); raw.kvc_debug_sel_trace(buf)
t = list(buf)
names = ["start", "loaded+pooled", "sums done", "minmax done", "thresholds done", "scan done", "compacted", "sorted", "written"]
for i in range(1, 9):
    print(f"{names[i]:16s} +{t[i] - t[i-1]:7d} cycles  (t={t[i] - t[0]})")
