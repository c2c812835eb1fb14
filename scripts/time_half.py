"""Scoring kernels alone on cfg-2 layer shapes: the full fused kernel on one stream, the half-footprint kernel on two.
usage: [KVC_STAGGER_NS=n] python scripts/time_half.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768; N = 16
g = torch.Generator(device=dev).manual_seed(3)
Ks = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(6)]
Q = torch.randn(1, 32, 32, 128, generator=g, device=dev).to(torch.bfloat16)
ss = [torch.cuda.Stream(priority=-1) for _ in range(2)]
def build(impl, nst):
    gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            for s in ss[:nst]: s.wait_stream(cap)
            keep = []
            for i in range(N):
                with torch.cuda.stream(ss[i % nst]):
                    keep.append(_native.window_scores(Q, Ks[i % 6], 32, "sum", impl, seq_len=L))
            for s in ss[:nst]: cap.wait_stream(s)
    torch.cuda.current_stream().wait_stream(cap)
    return gr, keep
for name, impl, nst in (("fused, 1 stream", _native.IMPL_TCGEN05_FUSED, 1), ("half, 2 streams", _native.IMPL_TCGEN05_HALF, 2)):
    _native.window_scores(Q, Ks[0], 32, "sum", impl, seq_len=L); torch.cuda.synchronize()
    gr, keep = build(impl, nst)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:18s} {e0.elapsed_time(e1) * 1e3 / 5 / N:6.2f} us per launch")
