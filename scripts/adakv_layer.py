"""One AdaKV cfg-4 layer (for ncu launch lists): update_kv on un-repeated K/V, then one flattened decode append."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import AdaKVCluster, _native
dev = torch.device("cuda:0")
B, Hq, Hkv, L, d, w, cap = 1, 32, 8, 65536, 128, 32, 2048
g = torch.Generator(device=dev).manual_seed(1)
K = torch.randn(B, Hkv, L, d, generator=g, device=dev).to(torch.bfloat16)
V = torch.randn(B, Hkv, L, d, generator=g, device=dev).to(torch.bfloat16)
Q = torch.randn(B, Hq, L, d, generator=g, device=dev).to(torch.bfloat16)
cl = AdaKVCluster(window_size=w, kernel_size=7, pooling="maxpool", max_capacity_prompt=cap, floor=0.2, normalize=True,
                  layer_idx=0, num_hidden_layers=32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    fk, fv = cl.update_kv(K, Q, V)
    state = torch.randn(Hq, d, generator=g, device=dev).to(torch.bfloat16)
    nk = _native.update_flatten_view(fk, state, cl.head_lens, cl.cu_klen)
torch.cuda.synchronize()
print("flat rows", fk.shape, "caps", cl.head_lens[:4].tolist())
