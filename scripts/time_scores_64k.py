import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native as N
dev = torch.device("cuda:0")
L, Hq, Hkv, d, w = 65536, 32, 8, 128, 32
gen = torch.Generator(device=dev); gen.manual_seed(1)
Ks = [torch.randn(1, Hkv, L, d, generator=gen, device=dev).to(torch.bfloat16) for _ in range(3)]
Q = torch.randn(1, Hq, w, d, generator=gen, device=dev).to(torch.bfloat16)
for name, impl in (("auto", N.IMPL_AUTO), ("fused", N.IMPL_TCGEN05_FUSED), ("dyn", N.IMPL_TCGEN05_DYN), ("half", N.IMPL_TCGEN05_HALF), ("tc", N.IMPL_TCGEN05)):
    for red in ("sum", "mean"):
        try:
            for i in range(3): N.window_scores(Q, Ks[i % 3], w, red, impl, seq_len=L)
        except Exception as e:
            print(name, red, "unsupported", str(e)[:80]); continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): N.window_scores(Q, Ks[i % 3], w, red, impl, seq_len=L)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:6s} {red:5s} {e0.elapsed_time(e1) * 1e3 / 20:7.1f} us")
