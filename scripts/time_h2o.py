"""H2O scoring alone on Mistral-7B shapes (32q / 8kv heads, d = 128): one (batch, layer) at L, both passes.
usage: [KVC_H2O_POLY=n] python scripts/time_h2o.py [L=16384] [B=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device=dev).manual_seed(1)
K = torch.randn(B, 8, L, 128, generator=g, device=dev).to(torch.bfloat16)
Q = torch.randn(B, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
for _ in range(2):
    s = _native.h2o_scores(Q, K, 32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    s = _native.h2o_scores(Q, K, 32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"L={L} B={B}: {ms:8.2f} ms  {2.0 * B * 32 * L * L / (ms * 1e-3) / 1e12:5.2f} Texp/s  checksum {float(s.float().sum()):.4f}")
