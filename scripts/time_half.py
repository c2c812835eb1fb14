"""Half-footprint scoring kernel: alone, and two launches in flight on two streams (the point of it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768; N = 16
g = torch.Generator(device=dev).manual_seed(7)
Ks = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(4)]
Q = torch.randn(1, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
ref = _native.window_scores(Q, Ks[0], 32, "sum", _native.IMPL_TCGEN05_FUSED)
out = _native.window_scores(Q, Ks[0], 32, "sum", _native.IMPL_TCGEN05_HALF)
torch.cuda.synchronize()
a, b = ref.view(torch.int16), out.view(torch.int16)
print(f"half vs fused: {100 * (a == b).float().mean().item():.3f}% bit-identical, max |ulp diff| {(a.int() - b.int()).abs().max().item()}")
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
def build(impl, streams, stagger_cycles=0):
    gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            for st in streams: st.wait_stream(cap)
            if stagger_cycles and len(streams) > 1:
                with torch.cuda.stream(streams[1]):
                    torch.cuda._sleep(stagger_cycles)     # second stream starts half a kernel later
            keep = []
            for i in range(N):
                with torch.cuda.stream(streams[i % len(streams)]):
                    keep.append(_native.window_scores(Q, Ks[i % 4], 32, "sum", impl))
            for st in streams: cap.wait_stream(st)
    torch.cuda.current_stream().wait_stream(cap)
    return gr, keep
def t(gr):
    gr.replay(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(5): gr.replay()
    y.record(); torch.cuda.synchronize()
    return x.elapsed_time(y) * 1e3 / 5 / N
for name, impl, streams in (("fused, 1 stream", _native.IMPL_TCGEN05_FUSED, [s0]), ("half, 1 stream", _native.IMPL_TCGEN05_HALF, [s0]),
                            ("half, 2 streams", _native.IMPL_TCGEN05_HALF, [s0, s1]),
                            ("half, 3 streams", _native.IMPL_TCGEN05_HALF, [s0, s1, torch.cuda.Stream()])):
    gr, keep = build(impl, streams)
    print(f"{name:18s} {t(gr):7.1f} us per layer")

for cyc in (10000, 20000, 30000, 40000, 60000):
    gr, keep = build(_native.IMPL_TCGEN05_HALF, [s0, s1], cyc)
    print(f"half, 2 streams, second delayed by {cyc} cycles: {t(gr):7.1f} us per layer")
