import os, sys
sys.path.insert(0, "/root/repo")
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768
Ks = [torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
Q = torch.randn(1, 32, L, 128, device=dev).to(torch.bfloat16)
for w in (8, 16, 32):
    gr = torch.cuda.CUDAGraph()
    for i in range(3): _native.window_scores(Q, Ks[i], w, "sum", 0)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        keep = [_native.window_scores(Q, Ks[i % 4], w, "sum", 0) for i in range(16)]
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): gr.replay()
    b.record(); torch.cuda.synchronize()
    print(f"w={w:2d}: {a.elapsed_time(b) * 1e3 / 80:6.1f} us per scoring launch")
