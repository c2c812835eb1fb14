"""One H2O scoring call at a reduced cfg-3 shape (for ncu): B=1, 32q/8kv heads, L=16384."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device=dev).manual_seed(1)
K = torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16)
Q = torch.randn(1, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
for _ in range(2):
    s = _native.h2o_scores(Q, K, 32)
torch.cuda.synchronize()
print(s.shape)
