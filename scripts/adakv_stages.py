"""Stage times of one AdaKV prefill layer at BASELINE configs[3] (Llama-3-8B, L = 65536, cap 2048, floor 0.2, maxpool 7):
CUDA events around each library call of the five-call path, inputs rotating over 3 buffer sets (> L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native as N
dev = torch.device("cuda:0")
L, Hq, Hkv, d, w, cap = 65536, 32, 8, 128, 32, 2048
gen = torch.Generator(device=dev); gen.manual_seed(4242)
Ks = [torch.randn(1, Hkv, L, d, generator=gen, device=dev).to(torch.bfloat16) for _ in range(3)]
Vs = [torch.randn(1, Hkv, L, d, generator=gen, device=dev).to(torch.bfloat16) for _ in range(3)]
Q = torch.zeros(1, L, Hq, d, device=dev, dtype=torch.bfloat16)
Q[:, L - w:] = torch.randn(1, w, Hq, d, generator=gen, device=dev).to(torch.bfloat16)
Q = Q.transpose(1, 2)
base, n = cap - w, L - w
names = ["scores", "select(sums)", "allocate", "read caps", "select(k per row)", "gather varlen"]
acc = [0.0] * len(names)
reps = int(os.environ.get("ADAKV_REPS", "20"))
for it in range(reps + 3):
    K, V = Ks[it % 3], Vs[it % 3]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    scores = N.window_scores(Q, K, w, "mean", N.IMPL_AUTO); ev[1].record()
    _, pooled, sums = N.pool_topk(scores, 7, "maxpool", base, want_pooled=True, want_sums=True, sums_only=True); ev[2].record()
    c, head_lens, cu_klen = N.adakv_allocate(pooled.view(Hq, n), sums, base, 0.2, True, w); ev[3].record()
    ch = c.cpu(); max_cap = int(ch.max()); klen = int(ch.sum()) + Hq * w; ev[4].record()
    idx, _, _ = N.pool_topk(scores, 7, "maxpool", max_cap, k_per_row=c); ev[5].record()
    fk, fv = N.gather_varlen(K, V, idx.view(Hq, max_cap), c, cu_klen, Hq, w, max_cap, klen); ev[6].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i in range(len(names)):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3
print(f"max_cap {max_cap}, rows {klen}")
for nm, a in zip(names, acc):
    print(f"{nm:20s} {a / reps:8.1f} us")
print(f"{'total':20s} {sum(acc) / reps:8.1f} us")
