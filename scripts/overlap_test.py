"""Does the gather (and select) kernel run concurrently with the scoring kernel?  Times N layers of
scoring on one stream and N gathers / selects on another, alone and together (CUDA graphs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768; N = 8
g = torch.Generator(device=dev).manual_seed(3)
Ks = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(N)]
Vs = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(N)]
Q = torch.randn(1, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
sc = [_native.window_scores(Q, Ks[i], 32, "sum", 0) for i in range(N)]
idx = [_native.pool_topk(sc[i], 7, "avgpool", 2016)[0] for i in range(N)]
torch.cuda.synchronize()
hi, lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)

def build(do_s, do_g, do_sel, sel_on_hi=False):
    gr = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            hi.wait_stream(cap); lo.wait_stream(cap)
            keep = []
            if do_s:
                with torch.cuda.stream(hi):
                    keep += [_native.window_scores(Q, Ks[i], 32, "sum", 0) for i in range(N)]
            if sel_on_hi and do_sel:
                with torch.cuda.stream(hi):
                    for i in range(N):
                        keep.append(_native.pool_topk(sc[i], 7, "avgpool", 2016)[0])
            with torch.cuda.stream(lo):
                for i in range(N):
                    if do_sel and not sel_on_hi: keep.append(_native.pool_topk(sc[i], 7, "avgpool", 2016)[0])
                    if do_g: keep.append(_native.gather_dense(Ks[i], Vs[i], idx[i], 32, 2016, 32))
            cap.wait_stream(hi); cap.wait_stream(lo)
    torch.cuda.current_stream().wait_stream(cap)
    return gr, keep

def t(gr):
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / 5 / N

for name, cfg in (("scores alone", (1, 0, 0)), ("gather alone", (0, 1, 0)), ("select alone", (0, 0, 1)),
                  ("scores || gather", (1, 1, 0)), ("scores || select", (1, 0, 1)), ("scores || select+gather", (1, 1, 1)),
                  ("select || gather", (0, 1, 1, True)), ("select || select", "selsel")):
    if cfg == "selsel":
        gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            with torch.cuda.graph(gr, stream=cap):
                hi.wait_stream(cap); lo.wait_stream(cap); keep = []
                for i in range(N):
                    with torch.cuda.stream(hi if i % 2 == 0 else lo):
                        keep.append(_native.pool_topk(sc[i], 7, "avgpool", 2016)[0])
                cap.wait_stream(hi); cap.wait_stream(lo)
        torch.cuda.current_stream().wait_stream(cap)
    else:
        gr, keep = build(*cfg)
    print(f"{name:26s} {t(gr):7.1f} us per layer")
