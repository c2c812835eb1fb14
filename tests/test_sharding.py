"""Multi-GPU host logic on CPU: (batch, kv-head) unit partition and the AdaKV budget exchange over a
world_size-2 gloo group.  The GPU half (histogram kernel, sharded cluster) is in
the gpu-marked tests at the bottom."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvcache_factory_b200 import sharding
from kvcache_factory_b200.synth import Shape, make_qkv
from oracle import kvc_oracle as O
from tests.helpers import to_bits


def _keys(t: torch.Tensor) -> np.ndarray:
    bits = to_bits(t).astype(np.uint32)
    return np.where(bits & 0x8000, (~bits) & 0xFFFF, bits | 0x8000).astype(np.int64)


def _hist(adaptive: torch.Tensor) -> torch.Tensor:
    """[H, 65536] histogram of the ordered 16-bit keys of a [H, n] score tensor (CPU restatement)."""
    k = _keys(adaptive)
    return torch.stack([torch.from_numpy(np.bincount(k[h], minlength=65536)) for h in range(k.shape[0])])


def _case(H=8, L=700, w=16, base=64, dtype=torch.bfloat16, floor=0.2, normalize=True):
    shape = Shape(1, H, 2, L, 128)
    Q, K, V = make_qkv(shape, dtype, layer_idx=3, dist="peaked", window=w)
    pooled = O.adakv_scores(K, Q, w, 7, "maxpool")
    cap_ref, _, adaptive = O.adakv_capacities(pooled, base, floor, normalize)
    return pooled, adaptive[0], cap_ref, base, floor


def _check_caps(cap, cap_ref, adaptive, H, base, floor):
    flat = adaptive.float().reshape(-1)
    kth = flat.topk(H * base).values[-1]
    ties = (adaptive.float() == kth).sum(-1).numpy()
    assert np.all(np.abs(cap.numpy() - cap_ref.numpy()) <= ties * (1 - floor) + 1), (cap, cap_ref, ties)
    assert abs(int(cap.sum()) - int(cap_ref.sum())) <= H
    if int((adaptive.float() == kth).sum()) == 1:      # no ties at the threshold: the budgets are exact
        assert torch.equal(cap, cap_ref)


def test_unit_partition_is_a_partition():
    for B, Hkv, world in [(1, 8, 1), (1, 8, 2), (1, 8, 8), (3, 8, 4), (5, 2, 4)]:
        seen = []
        for r in range(world):
            seen += sharding.unit_partition(B, Hkv, world, r)
        assert sorted(seen) == [(b, h) for b in range(B) for h in range(Hkv)]
    assert sharding.kv_head_slice(8, 4, 1) == slice(2, 4)
    with pytest.raises(ValueError):
        sharding.kv_head_slice(8, 3, 0)
    q, k = torch.zeros(1, 32, 5, 4), torch.zeros(1, 8, 5, 4)
    qs, ks, vs = sharding.shard_qkv(q, k, k, 4, 3)
    assert qs.shape[1] == 8 and ks.shape[1] == 2 and qs.data_ptr() == q[:, 24:].data_ptr()


@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_adakv_caps_single_process_matches_reference_logic(dtype, normalize):
    pooled, adaptive, cap_ref, base, floor = _case(dtype=dtype, normalize=normalize)
    H, n = adaptive.shape
    cap, cap_all, thr = sharding.adakv_caps_from_histograms(_hist(adaptive), H, base, floor, n)
    assert torch.equal(cap, cap_all) and cap.dtype == torch.int32
    _check_caps(cap, cap_ref, adaptive, H, base, floor)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, hist, H, base, floor, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = H // world
    cap, cap_all, thr = sharding.adakv_caps_from_histograms(hist[rank * per:(rank + 1) * per], H, base, floor, n,
                                                            None, rank, world)
    out[rank] = (cap.clone(), cap_all.clone(), thr)
    dist.barrier()
    dist.destroy_process_group()


def test_adakv_caps_world2_gloo_equals_single_process():
    pooled, adaptive, cap_ref, base, floor = _case()
    H, n = adaptive.shape
    hist = _hist(adaptive)
    single, _, thr1 = sharding.adakv_caps_from_histograms(hist, H, base, floor, n)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, hist, H, base, floor, n, out), nprocs=2, join=True)
    assert out[0][2] == out[1][2] == thr1
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][1], single)       # every rank sees all budgets
    assert torch.equal(torch.cat([out[0][0], out[1][0]]), single)                      # local slices tile them
    _check_caps(single, cap_ref, adaptive, H, base, floor)


# ---- GPU half ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [True, False])
def test_key_histogram_kernel_matches_cpu(normalize):
    from kvcache_factory_b200 import _native
    pooled, adaptive, cap_ref, base, floor = _case(normalize=normalize)
    H, n = adaptive.shape
    dev = torch.device("cuda:0")
    p = pooled[0].to(dev)
    _, _, sums = _native.pool_topk(p, 1, None, base, want_sums=True)
    hist = _native.adakv_key_histogram(p, sums, normalize)
    assert torch.equal(hist.cpu().long(), _hist(adaptive))


@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [True, False])
def test_two_level_histogram_kernels_match_cpu(normalize):
    """kvc_adakv_hist_coarse / kvc_adakv_hist_fine (what ShardedAdaKV exchanges) against the 65 536-bin CPU histogram,
    and the capacities they lead to against the full-resolution path."""
    from kvcache_factory_b200 import _native
    pooled, adaptive, cap_ref, base, floor = _case(normalize=normalize, L=9000, base=300)
    H, n = adaptive.shape
    dev = torch.device("cuda:0")
    p = pooled[0].to(dev)
    _, _, sums = _native.pool_topk(p, 1, None, base, want_sums=True)
    full = _hist(adaptive)
    akeys, h1 = _native.adakv_hist_coarse(p, sums, normalize)
    assert torch.equal(h1.cpu().long(), full.view(H, 4096, 16).sum(2))
    # (the oracle's normalised scores are sorted per head, as the reference sorts before normalising: compare as multisets)
    assert np.array_equal(np.sort(akeys.cpu().numpy().view(np.uint16).astype(np.int64), axis=1), np.sort(_keys(adaptive), axis=1))
    for b in (int(full.sum(0).view(4096, 16).sum(1).nonzero().max()), int(full.sum(0).view(4096, 16).sum(1).argmax())):
        h2, above = _native.adakv_hist_fine(akeys, h1, torch.tensor([b], dtype=torch.int32, device=dev))
        assert torch.equal(h2.cpu().long(), full[:, b * 16:(b + 1) * 16])
        assert torch.equal(above.cpu().long(), full[:, (b + 1) * 16:].sum(1))
    cap2, _, thr2 = sharding.adakv_caps_two_level(lambda: h1, lambda t: _native.adakv_hist_fine(akeys, h1, t), H, base, floor, n)
    cap1, _, thr1 = sharding.adakv_caps_from_histograms(full, H, base, floor, n)
    assert int(thr2) == thr1 and torch.equal(cap2.cpu(), cap1)


@pytest.mark.gpu
def test_sharded_adakv_world1_equals_cluster():
    from kvcache_factory_b200 import AdaKVCluster
    shape = Shape(1, 8, 2, 900, 128)
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=2, dist="peaked", window=32)
    dev = torch.device("cuda:0")
    a = AdaKVCluster(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=128, floor=0.2, normalize=True)
    b = AdaKVCluster(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=128, floor=0.2, normalize=True)
    ka, va = a.update_kv(K.to(dev), Q.to(dev), V.to(dev))
    kb, vb = sharding.ShardedAdaKV(b).update_kv(K.to(dev), Q.to(dev), V.to(dev))
    assert torch.equal(a.head_lens, b.head_lens) and torch.equal(a.cu_klen, b.cu_klen)
    assert torch.equal(ka, kb) and torch.equal(va, vb)
