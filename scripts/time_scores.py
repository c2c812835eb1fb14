"""Time the window-scoring kernel alone (cfg-2 layer shape) and compare implementations bit for bit.
usage: python scripts/time_scores.py [L=32768] [iters=20]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
Ks = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(4)]  # 4 x 64 MiB > L2
Q = torch.randn(1, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
outs = {}
for name, impl in (("tc", _native.IMPL_TCGEN05), ("fused", _native.IMPL_TCGEN05_FUSED)):
    for i in range(4):
        s = _native.window_scores(Q, Ks[i % 4], 32, "sum", impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        s = _native.window_scores(Q, Ks[i % 4], 32, "sum", impl)
    e1.record()
    torch.cuda.synchronize()
    outs[name] = _native.window_scores(Q, Ks[0], 32, "sum", impl).clone()
    eager = e0.elapsed_time(e1) * 1e3 / iters
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = [_native.window_scores(Q, Ks[i % 4], 32, "sum", impl) for i in range(16)]
    gr.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:6s} eager {eager:8.1f} us per call, CUDA graph {e0.elapsed_time(e1) * 1e3 / 80:8.1f} us per call (K rotates over 256 MiB)")
a, b = outs["tc"].view(torch.int16), outs["fused"].view(torch.int16)
same = (a == b).float().mean().item()
d = (a.int() - b.int()).abs().max().item()
print(f"fused vs tc: {100 * same:.3f}% bit-identical, max |ulp diff| {d}")
