"""Scoring kernel next to a co-resident, sync-heavy dummy with a small-select footprint: does either slow down?"""
import os, sys, ctypes
os.environ.setdefault("KVC_FUSED_COOP", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768; N = 8
d = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "ubench", "bin", "libdummy.so"))
d.dummy_launch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
g = torch.Generator(device=dev).manual_seed(3)
Ks = [torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(N)]
Q = torch.randn(1, 32, L, 128, generator=g, device=dev).to(torch.bfloat16)
sc = torch.rand(32, 32768, device=dev).to(torch.bfloat16)
out = torch.zeros(4, dtype=torch.int32, device=dev)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hi, lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
def build(do_s, do_d):
    gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(gr, stream=cap):
            hi.wait_stream(cap); lo.wait_stream(cap); keep = []
            if do_s:
                with torch.cuda.stream(hi):
                    keep += [_native.window_scores(Q, Ks[i], 32, "sum", 0) for i in range(N)]
            if do_d:
                with torch.cuda.stream(lo):
                    for i in range(N):
                        rc = d.dummy_launch(sc.data_ptr(), 32, 32768, rounds, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                        assert rc == 0, rc
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
for name, cfg in (("scores alone", (1, 0)), ("dummy alone", (0, 1)), ("scores || dummy", (1, 1))):
    gr, keep = build(*cfg)
    print(f"{name:20s} {t(gr):7.1f} us per layer")
