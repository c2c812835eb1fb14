"""Multi-GPU layout of the compression path: one process per GPU, KV-head shards.

Every score / select / gather is independent per (batch, query head) and shares K/V only inside a
GQA group, so the shardable unit is ``(batch, kv_head)`` with its ``Hq/Hkv`` query heads.  SnapKV,
PyramidKV, H2O and StreamingLLM need NO collective: each rank runs the ordinary cluster on the
units it owns.  AdaKV has exactly one exchange: the global top-(H*base) over all heads' normalised
scores (reference ``pyramidkv_utils.py:712-717``).  Scores are 16-bit floats, so that top-k is
fully determined by histograms of an order-preserving key, taken at the two levels the single-GPU
allocation kernels use (``adakv.cu``):

    local [H_local, 4096] histogram of key >> 4   (CUDA: kvc_adakv_hist_coarse)
      -> sum over local heads -> all_reduce(SUM)            16 KiB
      -> threshold bin (device scalar, never read by the host)
    local [H_local, 16] histogram inside that bin + per-head count above it   (kvc_adakv_hist_fine)
      -> sum over local heads -> all_reduce(SUM)            64 B -> threshold key T
      -> all_gather of per-head (above, equal) counts       2*H_local int64 per rank
      -> ties at T go to the lowest GLOBAL head index first  (torch.topk leaves tie order unspecified)
      -> cap_h = round(count_h*(1-floor) + int(base*floor))  same ATen ops as the reference (:719)

The host logic below is device-agnostic torch (tiny tensors), so it is tested on CPU with the gloo
backend at world_size 2; on GPUs the same code runs over NCCL / NVLink.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def unit_partition(batch: int, kv_heads: int, world_size: int, rank: int) -> List[Tuple[int, int]]:
    """(batch, kv_head) units owned by ``rank``: round-robin over the flattened unit index."""
    return [(u // kv_heads, u % kv_heads) for u in range(batch * kv_heads) if u % world_size == rank]


def kv_head_slice(kv_heads: int, world_size: int, rank: int) -> slice:
    """Contiguous KV-head shard of head-parallel tensor parallelism (kv_heads % world_size == 0)."""
    if kv_heads % world_size:
        raise ValueError(f"{kv_heads} KV heads do not split evenly over {world_size} ranks")
    per = kv_heads // world_size
    return slice(rank * per, (rank + 1) * per)


def shard_qkv(query_states, key_states, value_states, world_size: int, rank: int):
    """Views of Q/K/V restricted to this rank's KV heads and their query groups (no copy)."""
    Hq, Hkv = query_states.shape[1], key_states.shape[1]
    s = kv_head_slice(Hkv, world_size, rank)
    g = Hq // Hkv
    return (query_states[:, s.start * g: s.stop * g], key_states[:, s], value_states[:, s])


def adakv_caps_two_level(coarse, fine, total_heads: int, base_capacity: int, floor_ratio: float, row_len: int,
                         group: Optional["dist.ProcessGroup"] = None, rank: int = 0, world_size: int = 1):
    """Head capacities of this rank's heads from the two-level key histograms of ``adakv.cu``.

    ``coarse() -> int [H_local, 4096]`` (histogram of ``key >> 4``); ``fine(thr_bin) -> (int [H_local, 16],
    int [H_local])`` (histogram of ``key & 15`` inside bin ``thr_bin``, a one-element int32 tensor on the histograms'
    device, and the count above that bin).  Heads are laid out rank-major.  Collectives (world_size > 1): all-reduce
    of 4096 int32 (16 KiB), all-reduce of 16 int32, all-gather of 2 ints per head.  Nothing is read back to the host
    between the phases.  Returns (cap_local int32 [H_local], cap_all int32 [total_heads], threshold key as a tensor).
    """
    K = total_heads * base_capacity
    h1 = coarse()
    H_local = h1.shape[0]
    if H_local * world_size != total_heads:
        raise ValueError("heads must be split evenly over the ranks")
    dev = h1.device
    tot1 = h1.sum(dim=0, dtype=torch.int32)
    if world_size > 1:
        dist.all_reduce(tot1, op=dist.ReduceOp.SUM, group=group)
    Kt = torch.tensor([K], dtype=torch.int64, device=dev)
    desc1 = tot1.flip(0).to(torch.int64).cumsum(0)               # running count from the top bin down
    pos1 = torch.searchsorted(desc1, Kt).clamp(max=4095)         # first bin (from the top) reaching K
    thr_bin = (4095 - pos1).to(torch.int32)
    above_bin = desc1[pos1] - tot1.flip(0)[pos1].to(torch.int64)  # keys in the bins above it
    h2, above = fine(thr_bin)
    tot2 = h2.sum(dim=0, dtype=torch.int32)
    if world_size > 1:
        dist.all_reduce(tot2, op=dist.ReduceOp.SUM, group=group)
    desc2 = tot2.flip(0).to(torch.int64).cumsum(0) + above_bin
    pos2 = torch.searchsorted(desc2, Kt).clamp(max=15)
    thr_low = 15 - pos2                                          # low four bits of the threshold key
    lows = torch.arange(16, device=dev)
    h2l = h2.to(torch.int64)
    gt = above.to(torch.int64) + (h2l * (lows > thr_low)).sum(dim=1)
    eq = h2l.gather(1, thr_low.expand(H_local, 1)).squeeze(1)
    mine = torch.stack([gt, eq], dim=1).contiguous()
    if world_size > 1:
        parts = [torch.empty_like(mine) for _ in range(world_size)]
        dist.all_gather(parts, mine, group=group)
        allc = torch.cat(parts, dim=0)
    else:
        allc = mine
    above_all, equal_all = allc[:, 0], allc[:, 1]
    remaining = K - above_all.sum()
    before = equal_all.cumsum(0) - equal_all                      # ties claimed by lower global heads
    take = (remaining - before).clamp(min=0).minimum(equal_all)
    counts = above_all + take
    floor_capacity = int(base_capacity * floor_ratio)
    cap_all = torch.round(counts * (1 - floor_ratio) + floor_capacity).int().clamp(max=row_len)   # :719
    cap_local = cap_all[rank * H_local: (rank + 1) * H_local].contiguous()
    return cap_local, cap_all, (thr_bin.to(torch.int64) * 16 + thr_low).squeeze(0)


def adakv_caps_from_histograms(hist_local: torch.Tensor, total_heads: int, base_capacity: int, floor_ratio: float,
                               row_len: int, group: Optional["dist.ProcessGroup"] = None,
                               rank: int = 0, world_size: int = 1):
    """The same from full-resolution histograms (``int [H_local, 65536]``, e.g. ``kvc_adakv_key_histogram`` or a CPU
    bincount): the two levels are cut out of them.  Returns (cap_local, cap_all, threshold key as an int)."""
    H_local = hist_local.shape[0]

    def coarse():
        return hist_local.view(H_local, 4096, 16).sum(dim=2)

    def fine(thr_bin):
        b = int(thr_bin)
        return (hist_local[:, b * 16:(b + 1) * 16], hist_local[:, (b + 1) * 16:].sum(dim=1))

    cap_local, cap_all, thr = adakv_caps_two_level(coarse, fine, total_heads, base_capacity, floor_ratio, row_len,
                                                   group, rank, world_size)
    return cap_local, cap_all, int(thr)


class ShardedAdaKV:
    """AdaKV over a KV-head shard: same outputs as ``AdaKVCluster.update_kv`` for the local heads,
    with the head budgets decided globally over all ranks' heads."""

    def __init__(self, cluster, group: Optional["dist.ProcessGroup"] = None):
        self.cluster = cluster
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def update_kv(self, key_states, query_states, value_states):
        from . import _native as N
        c = self.cluster
        bsz, heads_local, q_len, head_dim = query_states.shape
        assert bsz == 1
        w, n = c.window_size, q_len - c.window_size
        device = key_states.device
        scores = N.window_scores(query_states, key_states, w, "mean", c.score_impl)
        idx_base, pooled, sums = N.pool_topk(scores, c.kernel_size, c.pooling, c.base_capacity,
                                             want_pooled=True, want_sums=True)
        keys = {}

        def coarse():
            keys["akeys"], keys["hist1"] = N.adakv_hist_coarse(pooled.view(heads_local, n), sums, bool(c.normalize))
            return keys["hist1"]

        def fine(thr_bin):
            return N.adakv_hist_fine(keys["akeys"], keys["hist1"], thr_bin)

        cap, cap_all, _ = adakv_caps_two_level(coarse, fine, heads_local * self.world, c.base_capacity, c.floor_ratio,
                                               n, self.group, self.rank, self.world)
        cap = cap.to(device=device, dtype=torch.int32)
        head_lens = cap + w
        cu_klen = torch.zeros(heads_local + 1, dtype=torch.int32, device=device)
        cu_klen[1:] = head_lens.cumsum(0)
        max_cap = int(cap.max())
        klen_sum = int(head_lens.sum())
        if max_cap <= c.base_capacity:
            idx = idx_base.view(heads_local, c.base_capacity)
        else:
            idx, _, _ = N.pool_topk(scores, c.kernel_size, c.pooling, max_cap, k_per_row=cap)
            idx = idx.view(heads_local, max_cap)
        flat_k, flat_v = N.gather_varlen(key_states, value_states, idx, cap, cu_klen, heads_local, w, max_cap, klen_sum)
        c._init_metadata(device, heads_local, head_lens, cu_klen, klen_sum, max_cap + w)
        self.cap_all = cap_all
        return flat_k, flat_v
