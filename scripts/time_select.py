"""Select phase alone: 32 layers of cfg-2 rows (32 x 32736 scores, PyramidKV budgets), one stream vs rotating streams.
usage: [KVC_SELECT_SINGLE=1] python scripts/time_select.py [nstreams=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native, pyramid_budget
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0"); n = 32736; nl = 32
g = torch.Generator(device=dev).manual_seed(1)
scores = [(torch.rand(32, n, generator=g, device=dev) ** 8).to(torch.bfloat16) for _ in range(nl)]
ks = [pyramid_budget(32768, 32, 2048, 32, l) for l in range(nl)]
ref = [_native.pool_topk(scores[l], 7, "avgpool", ks[l])[0] for l in (0, 31)]
for nstreams in (1, ns):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            for s in streams: s.wait_stream(cap)
            keep = []
            for l in range(nl):
                with torch.cuda.stream(streams[l % nstreams]):
                    keep.append(_native.pool_topk(scores[l], 7, "avgpool", ks[l])[0])
            for s in streams: cap.wait_stream(s)
    torch.cuda.current_stream().wait_stream(cap)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{nstreams} stream(s): {e0.elapsed_time(e1) * 1e3 / 5 / nl:6.2f} us per layer ({e0.elapsed_time(e1) / 5:.3f} ms per 32 layers)",
          "same as eager:", torch.equal(keep[0], ref[0]) and torch.equal(keep[31], ref[1]))
