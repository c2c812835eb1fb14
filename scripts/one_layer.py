"""Run a few PyramidKV update_kv calls of one cfg-2 layer (for ncu / quick timing).
usage: python scripts/one_layer.py [impl=auto|mma|tc] [iters=3] [layer=0] [L=32768]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import PyramidKVCluster, _native

impl = sys.argv[1] if len(sys.argv) > 1 else "auto"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layer = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
dev = torch.device("cuda:0")
B, Hq, Hkv, d, w = 1, 32, 8, 128, 32
g = torch.Generator(device=dev).manual_seed(1234 + layer)
K = torch.randn(B, Hkv, L, d, generator=g, device=dev).to(torch.bfloat16)
V = torch.randn(B, Hkv, L, d, generator=g, device=dev).to(torch.bfloat16)
Q = torch.zeros(B, L, Hq, d, device=dev, dtype=torch.bfloat16)
Q[:, L - w:] = torch.randn(B, w, Hq, d, generator=g, device=dev).to(torch.bfloat16)
Q = Q.transpose(1, 2)
cl = PyramidKVCluster(num_hidden_layers=32, window_size=w, max_capacity_prompt=2048, kernel_size=7, pooling="avgpool", layer_idx=layer)
cl.score_impl = {"auto": 0, "mma": 1, "tc": 2, "fused": 3}[impl]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    cl.update_kv(K, Q, V, None, 4)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Kc, Vc = cl.update_kv(K, Q, V, None, 4)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("impl", impl, "layer", layer, "us per update_kv (cold L2):", [round(t, 1) for t in ts], "out", tuple(Kc.shape))
