"""Kernel timeline of one CUDA-graph replay of the cfg-2 step (library built with `make TRACE=1`):
which kernels overlap, and where the gaps are.  usage: python scripts/timeline.py [layers=8] [--sync-tail]"""
import os, sys, ctypes
os.environ.setdefault("KVC_FUSED_COOP", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import PyramidKVCluster, _native, wait_for_compression
nl = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
sync_tail = "--sync-tail" in sys.argv
dev = torch.device("cuda:0"); L, w = 32768, 32
g = torch.Generator(device=dev)
Ks, Vs, Qs, cls = [], [], [], []
for l in range(nl):
    g.manual_seed(l)
    Ks.append(torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16))
    Vs.append(torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16))
    Qs.append(torch.randn(1, 32, w, 128, generator=g, device=dev).to(torch.bfloat16))
    c = PyramidKVCluster(num_hidden_layers=32, window_size=w, max_capacity_prompt=2048, kernel_size=7, pooling="avgpool", layer_idx=l * (32 // nl))
    c.async_tail = not sync_tail
    c.async_scores = not sync_tail and '--sync-scores' not in sys.argv
    cls.append(c)
def step():
    for l in range(nl):
        cls[l].update_kv_window(Ks[l], Qs[l], Vs[l], 4, L)
    wait_for_compression(dev)
step(); torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH)
buf = (ctypes.c_longlong * (3 * 4096))()
gr = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(priority=-1); cap.wait_stream(torch.cuda.current_stream())
raw.kvc_debug_timeline(buf, 0, 1)
with torch.cuda.stream(cap):
    with torch.cuda.graph(gr, stream=cap):
        step()
torch.cuda.current_stream().wait_stream(cap)
n0 = raw.kvc_debug_timeline(buf, 4096, 1)   # entries allocated so far (warm-up + capture)
gr.replay(); torch.cuda.synchronize()
raw.kvc_debug_timeline(buf, 0, 1)
gr.replay(); torch.cuda.synchronize()
n = raw.kvc_debug_timeline(buf, 4096, 0)
ev = [(buf[3 * i + 1], buf[3 * i + 2], buf[3 * i]) for i in range(n) if 0 < buf[3 * i + 1] < (1 << 62) and buf[3 * i + 2] > 0]
ev.sort()
t0 = ev[0][0]
names = {0: "SCORE ", 1: "select", 2: "gather"}
print(f"{len(ev)} kernels in the replay; times in us from the first kernel start")
for s, e, k in ev:
    print(f"  {names[k]}  {(s - t0) / 1e3:8.1f} -> {(e - t0) / 1e3:8.1f}  ({(e - s) / 1e3:6.1f})")
print("total", (max(e for _, e, _ in ev) - t0) / 1e3, "us")
