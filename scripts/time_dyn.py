"""Time the scoring kernels alone on one cfg-2 layer shape (or a 70B shard) and compare them bit for bit.
usage: python scripts/time_dyn.py [L=32768] [Hq=32] [Hkv=8] [streams=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
Hq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Hkv = int(sys.argv[3]) if len(sys.argv) > 3 else 8
nstreams = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
nbuf = max(4, (300 << 20) // (Hkv * L * 256) + 1)                     # rotate over > 2 x L2
Ks = [torch.randn(1, Hkv, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
Q = torch.randn(1, Hq, 32, 128, generator=g, device=dev).to(torch.bfloat16)
impls = [("fused", _native.IMPL_TCGEN05_FUSED), ("dyn", _native.IMPL_TCGEN05_DYN)]
outs = {}
streams = [torch.cuda.Stream() for _ in range(nstreams)]
for name, impl in impls:
    try:
        outs[name] = _native.window_scores(Q, Ks[0], 32, "sum", impl, seq_len=L).clone()
    except _native.KvcError as e:
        print(name, "unsupported:", e); continue
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    n = 32
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            keep = []
            if nstreams > 1:
                for st in streams: st.wait_stream(cap)
            for i in range(n):
                if nstreams > 1:
                    with torch.cuda.stream(streams[i % nstreams]):
                        keep.append(_native.window_scores(Q, Ks[i % nbuf], 32, "sum", impl, seq_len=L))
                else:
                    keep.append(_native.window_scores(Q, Ks[i % nbuf], 32, "sum", impl, seq_len=L))
            if nstreams > 1:
                for st in streams: cap.wait_stream(st)
    torch.cuda.current_stream().wait_stream(cap)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * n)
    by = Hkv * L * 256 + Hq * 32 * 256
    print(f"{name:6s} {nstreams} stream(s): {us:7.2f} us per launch  = {by / us / 1e3:7.1f} GB/s of algorithmic bytes ({by / us / 1e3 / 6534.8:.3f} of measured HBM peak)")
    same = torch.equal(keep[0], outs[name])
    print(f"       graph replay == eager: {same}")
if "fused" in outs and "dyn" in outs:
    a, b = outs["fused"].view(torch.int16), outs["dyn"].view(torch.int16)
    print(f"dyn vs fused: {100 * (a == b).float().mean().item():.3f}% bit-identical, max |ulp diff| {(a.int() - b.int()).abs().max().item()}")
