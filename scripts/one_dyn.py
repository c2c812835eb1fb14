"""A few launches of one scoring kernel on the cfg-2 layer shape (for ncu).
usage: python scripts/one_dyn.py [L=32768] [Hq=32] [Hkv=8] [impl=dyn|half|fused]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
Hq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Hkv = int(sys.argv[3]) if len(sys.argv) > 3 else 8
impl = {"dyn": _native.IMPL_TCGEN05_DYN, "half": _native.IMPL_TCGEN05_HALF, "fused": _native.IMPL_TCGEN05_FUSED}[sys.argv[4] if len(sys.argv) > 4 else "dyn"]
dev = torch.device("cuda:0")
Ks = [torch.randn(1, Hkv, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
Q = torch.randn(1, Hq, 32, 128, device=dev).to(torch.bfloat16)
for i in range(6):
    s = _native.window_scores(Q, Ks[i % 4], 32, "sum", impl, seq_len=L)
torch.cuda.synchronize()
print("ok", float(s.float().sum()))
