"""One cfg-5 shard layer (1 KV head / 8 query heads of Llama-3-70B, L=131072) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import PyramidKVCluster
dev = torch.device("cuda:0"); L, w = 131072, 32
g = torch.Generator(device=dev).manual_seed(1)
K = torch.randn(1, 1, L, 128, generator=g, device=dev).to(torch.bfloat16)
V = torch.randn(1, 1, L, 128, generator=g, device=dev).to(torch.bfloat16)
Qw = torch.randn(1, 8, w, 128, generator=g, device=dev).to(torch.bfloat16)
cl = PyramidKVCluster(num_hidden_layers=80, window_size=w, max_capacity_prompt=2048, kernel_size=7, pooling="avgpool", layer_idx=40)
for _ in range(3):
    o = cl.update_kv_window(K, Qw, V, 8, L)
torch.cuda.synchronize(); print(o[0].shape)
