import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import PyramidKVCluster, wait_for_compression, _native
dev = torch.device("cuda:0"); L, w, cap, nl = 32768, 32, 2048, 32
g = torch.Generator(device=dev)
Ks, Vs, Qs = [], [], []
for l in range(8):
    g.manual_seed(l)
    Ks.append(torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16))
    Vs.append(torch.randn(1, 8, L, 128, generator=g, device=dev).to(torch.bfloat16))
    Qs.append(torch.randn(1, 32, w, 128, generator=g, device=dev).to(torch.bfloat16))
def run(async_mode, impl):
    outs = []
    for l in range(nl):
        cl = PyramidKVCluster(num_hidden_layers=nl, window_size=w, max_capacity_prompt=cap, kernel_size=7, pooling="avgpool", layer_idx=l)
        cl.score_impl = impl
        cl.async_tail = async_mode; cl.async_scores = async_mode
        outs.append(cl.update_kv_window(Ks[l % 8], Qs[l % 8], Vs[l % 8], 4, L))
    wait_for_compression(dev); torch.cuda.synchronize()
    return outs
ref = run(False, _native.IMPL_TCGEN05_HALF)
bad = 0
for rep in range(30):
    out = run(True, _native.IMPL_AUTO)
    for l in range(nl):
        if not (torch.equal(out[l][0], ref[l][0]) and torch.equal(out[l][1], ref[l][1])):
            bad += 1
    del out
print("mismatching (layer, rep) pairs over 30 async steps of 32 layers:", bad)
