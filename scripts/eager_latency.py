"""Host-side cost of the drop-in default: eager update_kv per layer, fused one-call path vs three calls (cfg-1 and cfg-2 shapes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import SnapKVCluster, PyramidKVCluster
dev = torch.device("cuda:0")
for name, L, cap, mk in (("cfg-1 SnapKV L=4096 cap=128", 4096, 128, lambda: SnapKVCluster(32, 128, 7, "avgpool")),
                         ("cfg-2 PyramidKV layer 16 L=32768", 32768, 2048,
                          lambda: PyramidKVCluster(num_hidden_layers=32, window_size=32, max_capacity_prompt=2048, kernel_size=7, pooling="avgpool", layer_idx=16))):
    Ks = [torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
    Vs = [torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
    Q = torch.randn(1, L, 32, 128, device=dev).to(torch.bfloat16).transpose(1, 2)
    for fused in (True, False):
        cl = mk(); cl.fused_call = fused
        for i in range(8): cl.update_kv(Ks[i % 4], Q, Vs[i % 4], None, 4)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); n = 200
        for i in range(n): cl.update_kv(Ks[i % 4], Q, Vs[i % 4], None, 4)
        host = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e6
        print(f"{name:36s} {'one call ' if fused else 'three calls'}: host {host:6.1f} us per update_kv, wall {wall:6.1f} us")
