"""The asynchronous path (two half-footprint scoring launches in flight, select and gather on side streams) must give the
synchronous result bit for bit, step after step: no deadlock between the barrier grids, no cross-launch interference.
Shapes: a cfg-2 layer set (one row block per KV head) and the 70B shard (two row blocks per KV head, each a unit)."""
import pytest
import torch

from kvcache_factory_b200 import PyramidKVCluster, _native, wait_for_compression

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Hq,Hkv,L,layers_total", [(32, 8, 32768, 32), (16, 2, 131072, 80)])
def test_async_steps_equal_synchronous_run(Hq, Hkv, L, layers_total):
    dev = torch.device("cuda:0")
    w, cap, nl, nbuf = 32, 2048, 12, 3
    g = torch.Generator(device=dev)
    Ks, Vs, Qs = [], [], []
    for i in range(nbuf):
        g.manual_seed(100 + i)
        Ks.append(torch.randn(1, Hkv, L, 128, generator=g, device=dev).to(torch.bfloat16))
        Vs.append(torch.randn(1, Hkv, L, 128, generator=g, device=dev).to(torch.bfloat16))
        Qs.append(torch.randn(1, Hq, w, 128, generator=g, device=dev).to(torch.bfloat16))

    def run(async_mode, impl):
        outs = []
        for l in range(nl):
            cl = PyramidKVCluster(num_hidden_layers=layers_total, window_size=w, max_capacity_prompt=cap, kernel_size=7,
                                  pooling="avgpool", layer_idx=(l * layers_total) // nl)
            cl.score_impl = impl
            cl.async_tail = cl.async_scores = async_mode
            outs.append(cl.update_kv_window(Ks[l % nbuf], Qs[l % nbuf], Vs[l % nbuf], Hq // Hkv, L))
        wait_for_compression(dev)
        torch.cuda.synchronize()
        return outs

    ref = run(False, _native.IMPL_TCGEN05_HALF)        # the same kernel, one launch at a time
    for rep in range(8):
        out = run(True, _native.IMPL_AUTO)
        for l in range(nl):
            assert torch.equal(out[l][0], ref[l][0]) and torch.equal(out[l][1], ref[l][1]), (rep, l)
        del out
