import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0")
s = (torch.rand(32, 32736, device=dev) ** 4).to(torch.bfloat16)
for _ in range(3):
    idx, _, _ = _native.pool_topk(s, 7, "avgpool", 3932)
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH); buf = (ctypes.c_longlong * 64)(); raw.kvc_debug_seltrace(buf)
t = list(buf); names = ["start", "pooled", "minmax xchg done", "round1 counted", "r1 cluster sync", "round2 counted", "r2 cluster sync", "threshold known", "local scan done", "totals sync", "scatter done", "scatter sync", "sort+emit done (cta0)", "final sync"]
for i, nm in enumerate(names):
    print(f"{nm:26s} {t[i]-t[0]:8d}  (+{t[i]-t[i-1] if i else 0})")
