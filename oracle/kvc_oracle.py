"""CPU ORACLE for the prefill KV-compression hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (PyTorch CPU ATen ops + numpy integer work) of the algorithms in
the reference's ``pyramidkv/pyramidkv_utils.py`` and ``csrc/csrc/cuda_api.cu``.  It exists so the
CUDA path can be checked; it is NEVER on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it.  The product package (``kvcache_factory_b200``) has no CPU fallback and raises when
its CUDA library is missing.

Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the build container, runs it on seeded
inputs and commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays the
oracle against those vectors (and, where ``/root/reference`` is present, against the live
reference).  All arithmetic on the path is PyTorch ATen (torch unpinned by the reference's
``requirements.txt:5``; built here against torch 2.11.0), which is why the restatement uses the
same ATen CPU ops in the same order and dtype: the rounding points ARE the algorithm.

Differences from the reference that are deliberate and behaviour-preserving:
  * K and V may be passed un-repeated (``[B,Hkv,L,d]``); they are expanded with ``repeat_kv``
    exactly as every reference call site does before ``update_kv`` (``llama_model.py:379-380``).
  * Functions return the intermediate pooled scores and the selected indices next to the
    compressed tensors so tests can apply a tie-aware comparison.
  * No ``print`` per call.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# a1  GQA expansion
# --------------------------------------------------------------------------------------------
def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    """[B,Hkv,L,d] -> [B,Hkv*n_rep,L,d], head h of the output reads KV head h // n_rep.

    Follows ``pyramidkv_utils.py:108-117`` (equivalent to ``repeat_interleave(dim=1)``).
    """
    if n_rep == 1:
        return x
    B, Hkv, L, d = x.shape
    return x.unsqueeze(2).expand(B, Hkv, n_rep, L, d).reshape(B, Hkv * n_rep, L, d)


def _expand(K: torch.Tensor, V: torch.Tensor, Hq: int) -> Tuple[torch.Tensor, torch.Tensor]:
    Hkv = K.shape[1]
    assert Hq % Hkv == 0, "query heads must be a multiple of KV heads"
    return repeat_kv(K, Hq // Hkv), repeat_kv(V, Hq // Hkv)


# --------------------------------------------------------------------------------------------
# scoring (shared by SnapKV / PyramidKV / AdaKV / H2O)
# --------------------------------------------------------------------------------------------
def _block_causal_bias(w: int, logits_dtype: torch.dtype) -> torch.Tensor:
    """w x w additive bias: 0 on/below the diagonal, finfo(dtype).min above it.

    ``pyramidkv_utils.py:318-322``: the tensor is created with the default dtype (fp32) but filled
    with the *logits dtype's* minimum; adding it in place to bf16/fp16 logits rounds the sum back
    to that dtype.
    """
    bias = torch.full((w, w), torch.finfo(logits_dtype).min)
    col = torch.arange(w)
    bias.masked_fill_(col < (col + 1).view(w, 1), 0)
    return bias


def attention_probs(Kr: torch.Tensor, Q: torch.Tensor, w: int, all_queries: bool) -> torch.Tensor:
    """softmax(Q' K^T / sqrt(d)) rounded to the model dtype.

    Q' is the last ``w`` query rows (``pyramidkv_utils.py:317``) or, for H2O, every row (``:544``).
    Only the bottom-right w x w block is causally masked (``:324`` / ``:551``); everything else is
    unmasked, and the ``attention_mask`` argument of ``update_kv`` is ignored by the reference.
    Rounding points: GEMM output in model dtype, division result in model dtype, softmax computed
    in fp32 and cast back (``:326``).
    """
    d = Q.shape[-1]
    Qs = Q if all_queries else Q[..., -w:, :]
    logits = torch.matmul(Qs, Kr.transpose(2, 3)) / math.sqrt(d)
    logits[:, :, -w:, -w:] += _block_causal_bias(w, logits.dtype)[None, None]
    return F.softmax(logits, dim=-1, dtype=torch.float32).to(Q.dtype)


def window_scores(Kr: torch.Tensor, Q: torch.Tensor, w: int, reduce: str = "sum",
                  all_queries: bool = False) -> torch.Tensor:
    """Per-key importance [B,Hq,L-w] in model dtype.

    sum over the scored query rows (SnapKV ``:327``, PyramidKV ``:232/260``, H2O ``:554``) or mean
    (AdaKV ``:661``); the last ``w`` keys (the observation window itself) are excluded.
    """
    probs = attention_probs(Kr, Q, w, all_queries)
    rows = probs if all_queries else probs[:, :, -w:, :]
    rows = rows[..., :-w]
    if reduce == "sum":
        return rows.sum(dim=-2)
    if reduce == "mean":
        return rows.mean(dim=-2)
    raise ValueError(reduce)


def pool_scores(scores: torch.Tensor, kernel_size: int, pooling: Optional[str]) -> torch.Tensor:
    """stride-1 'same' pooling along the key axis (``pyramidkv_utils.py:328-333``).

    avg: zero padding, divisor always ``kernel_size``; max: -inf padding.  ``None`` = H2O (pooling
    commented out at ``:555-561``).  Only odd kernels keep the length (SURVEY.md §7 hard parts).
    """
    if pooling is None:
        return scores
    if kernel_size % 2 != 1:
        raise ValueError("kernel_size must be odd")
    if pooling == "avgpool":
        return F.avg_pool1d(scores, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)
    if pooling == "maxpool":
        return F.max_pool1d(scores, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)
    raise ValueError("Pooling method not supported")


def gather_and_append_window(Kr, Vr, idx: torch.Tensor, w: int):
    """K' = cat(K[:, :, :-w].gather(idx), K[:, :, -w:]) and the same for V (``:341-346``)."""
    d = Kr.shape[-1]
    gi = idx.unsqueeze(-1).expand(-1, -1, -1, d)
    Kc = torch.cat([Kr[:, :, :-w, :].gather(2, gi), Kr[:, :, -w:, :]], dim=2)
    Vc = torch.cat([Vr[:, :, :-w, :].gather(2, gi), Vr[:, :, -w:, :]], dim=2)
    return Kc, Vc


# --------------------------------------------------------------------------------------------
# a2  SnapKV, a3 PyramidKV, a4 H2O, a5 StreamingLLM
# --------------------------------------------------------------------------------------------
def _select(K, Q, V, w, n_keep, kernel_size, pooling, reduce="sum", all_queries=False):
    Kr, Vr = _expand(K, V, Q.shape[1])
    pooled = pool_scores(window_scores(Kr, Q, w, reduce, all_queries), kernel_size, pooling)
    idx = pooled.topk(n_keep, dim=-1).indices
    Kc, Vc = gather_and_append_window(Kr, Vr, idx, w)
    return Kc, Vc, idx, pooled


def snapkv_update_kv(K, Q, V, window_size=64, max_capacity_prompt=320, kernel_size=5,
                     pooling="avgpool"):
    """``SnapKVCluster.update_kv`` (``pyramidkv_utils.py:306-347``).

    Returns (K', V', idx, pooled); idx/pooled are None on the pass-through branch (``:314-315``),
    where K', V' are the (expanded) inputs.
    """
    assert K.shape[-2] == Q.shape[-2]
    L = Q.shape[-2]
    if L < max_capacity_prompt:
        Kr, Vr = _expand(K, V, Q.shape[1])
        return Kr, Vr, None, None
    return _select(K, Q, V, window_size, max_capacity_prompt - window_size, kernel_size, pooling)


def think_key_pruner(kv_states, q_states, recent_size=128, ratio=0.3):
    """``key_pruner_query_driven`` (``pyramidkv_utils.py:13-26``): drop the ``int(d*ratio)`` key channels with the smallest
    mean(q^2) * mean(k^2) from all but the last ``recent_size`` rows.  Returns (kv_pruned [1, B*H, rows-recent, d-k],
    kv_recent, kept-channel mask [B,H,d], channel scores [B,H,d])."""
    _, _, seqlen, head_dim = kv_states.shape
    k = int(head_dim * ratio)
    queries_norm = torch.pow(q_states[..., -32:, :], 2).mean(dim=2)      # the reference hard-codes 32 (:17)
    keys_norm = torch.pow(kv_states, 2).mean(dim=2)
    key = queries_norm * keys_norm
    _, indices = torch.topk(key, k, dim=-1, largest=False)
    keep_idx = indices.sort().values
    mask = torch.zeros(key.shape, dtype=torch.bool)
    mask = mask.scatter_(-1, keep_idx, 1)
    mask_k = mask.unsqueeze(2).expand(-1, -1, seqlen - recent_size, -1)
    pruned = kv_states[:, :, :seqlen - recent_size, :][~mask_k].reshape(1, -1, seqlen - recent_size, head_dim - k)
    return pruned, kv_states[:, :, seqlen - recent_size:, :], ~mask, key


def think_update_kv(K, Q, V, window_size=64, max_capacity_prompt=320, kernel_size=5, pooling="avgpool",
                    recent_size=32, ratio=0.4):
    """``SnapKVCluster.update_think`` (``pyramidkv_utils.py:349-392``): SnapKV compression, then channel pruning of the
    compressed keys.  Returns (kv_pruned, kv_recent, mask, V', K', channel scores), or (K, V) on the pass-through
    branch (``:357-358``)."""
    Kc, Vc, idx, _ = snapkv_update_kv(K, Q, V, window_size, max_capacity_prompt, kernel_size, pooling)
    if idx is None:
        return Kc, Vc
    pruned, recent, mask, key = think_key_pruner(Kc, Q, recent_size, ratio)
    return pruned, recent, mask, Vc, Kc, key


def pyramidkv_budget(q_len: int, window_size: int, max_capacity_prompt: int, num_hidden_layers: int,
                     layer_idx: int, beta: int = 20) -> int:
    """Per-layer number of non-window keys kept by PyramidKV (``pyramidkv_utils.py:205-215``)."""
    base = max_capacity_prompt - window_size
    min_num = base // beta
    max_num = base * 2 - min_num
    if max_num >= q_len - window_size:
        max_num = q_len - window_size
        min_num = base * 2 - max_num
    steps = (max_num - min_num) // (num_hidden_layers - 1)
    return max_num - layer_idx * steps


def pyramidkv_update_kv(K, Q, V, layer_idx, num_hidden_layers=32, window_size=64,
                        max_capacity_prompt=320, kernel_size=5, pooling="avgpool", beta=20):
    """``PyramidKVCluster.update_kv`` (``pyramidkv_utils.py:197-283``), three branches."""
    assert K.shape[-2] == Q.shape[-2]
    L = Q.shape[-2]
    base = max_capacity_prompt - window_size
    if L < max_capacity_prompt:                                   # :218-219
        Kr, Vr = _expand(K, V, Q.shape[1])
        return Kr, Vr, None, None
    if L < base * 2:                                              # :220-251 SnapKV budget
        n_keep = base
    else:                                                         # :252-283 pyramid budget
        n_keep = pyramidkv_budget(L, window_size, max_capacity_prompt, num_hidden_layers,
                                  layer_idx, beta)
    return _select(K, Q, V, window_size, n_keep, kernel_size, pooling)


def h2o_update_kv(K, Q, V, window_size=64, max_capacity_prompt=320):
    """``H2OKVCluster.update_kv`` (``pyramidkv_utils.py:533-575``): all query rows, no pooling."""
    assert K.shape[-2] == Q.shape[-2]
    L = Q.shape[-2]
    if L < max_capacity_prompt:
        Kr, Vr = _expand(K, V, Q.shape[1])
        return Kr, Vr, None, None
    return _select(K, Q, V, window_size, max_capacity_prompt - window_size, 1, None,
                   reduce="sum", all_queries=True)


def streamingllm_update_kv(K, Q, V, window_size=64, max_capacity_prompt=320):
    """``StreamingLLMKVCluster.update_kv`` (``:595-620``): first cap-w rows + last w rows."""
    assert K.shape[-2] == Q.shape[-2]
    B, Hq, L, d = Q.shape
    Kr, Vr = _expand(K, V, Hq)
    if L < max_capacity_prompt:
        return Kr, Vr, None, None
    n_keep = max_capacity_prompt - window_size
    idx = torch.arange(n_keep, dtype=torch.int64).view(1, 1, -1).expand(B, Hq, -1)
    Kc, Vc = gather_and_append_window(Kr, Vr, idx, window_size)
    return Kc, Vc, idx, None


def l2norm_update_kv(K, Q, V, max_capacity_prompt=320, layer_idx=0, skip_layers=()):
    """``L2NormCluster.update_kv`` (``pyramidkv_utils.py:406-429``): the ``max_capacity_prompt`` keys of smallest L2
    norm per head, ascending-norm order.  Returns (K', V', idx, norms); idx/norms None on the pass-through branches."""
    assert K.shape[-2] == Q.shape[-2]
    Kr, Vr = _expand(K, V, Q.shape[1])
    if Q.shape[-2] < max_capacity_prompt or layer_idx in skip_layers:
        return Kr, Vr, None, None
    d = Kr.size(-1)
    norms = torch.norm(Kr, p=2, dim=-1)                          # model dtype (:420)
    order = norms.argsort(dim=-1)                                # ascending (:421)
    gi = order.unsqueeze(-1).expand(-1, -1, -1, d)
    Kc = Kr.gather(2, gi)[:, :, :max_capacity_prompt, :]
    Vc = Vr.gather(2, gi)[:, :, :max_capacity_prompt, :]
    return Kc, Vc, order[..., :max_capacity_prompt], norms


# --------------------------------------------------------------------------------------------
# a6  AdaKV
# --------------------------------------------------------------------------------------------
def varlen_metadata(head_lens: List[int]) -> Dict[str, object]:
    """Flattened-cache metadata built by ``AdaKVCluster.update_kv.init_metadata`` (``:682-698``)."""
    H = len(head_lens)
    hl = np.asarray(head_lens, dtype=np.int32)
    cu_headlens = np.cumsum(hl, dtype=np.int32)
    klen_sum = int(hl.sum())
    cu_klen = np.concatenate([cu_headlens - hl, np.asarray([klen_sum], dtype=np.int32)])
    return dict(
        head_lens=hl,
        cu_headlens=cu_headlens,
        cu_klen=cu_klen.astype(np.int32),
        layer_qlens=np.ones(H, dtype=np.int32),
        cu_qlen=np.arange(0, H + 1, dtype=np.int32),
        cu_offset=np.arange(0, H + 1, dtype=np.int32),
        cu_head_offset=np.arange(1, H + 1, dtype=np.int32),
        klen_sum=klen_sum,
        qlen_sum=H,
        max_seqlen_k=int(hl.max()) if H else 0,
    )


def adakv_scores(K, Q, window_size=32, kernel_size=7, pooling="maxpool") -> torch.Tensor:
    """``AdaKVCluster.calcul_attn_sore`` (``:647-672``): window MEAN, then pooling."""
    Kr = repeat_kv(K, Q.shape[1] // K.shape[1])
    return pool_scores(window_scores(Kr, Q, window_size, "mean"), kernel_size, pooling)


def adakv_capacities(pooled: torch.Tensor, base_capacity: int, floor_ratio: float,
                     normalize: bool):
    """Head-adaptive budgets (``pyramidkv_utils.py:705-719``).

    Returns (cap_h int32 [H], sorted_idx int64 [1,H,n], normalised sorted scores [1,H,n]).
    """
    B, H, n = pooled.shape
    floor_capacity = int(base_capacity * floor_ratio)
    sorted_score, sorted_idx = pooled.sort(dim=-1, descending=True)
    adaptive = sorted_score
    if normalize:
        ratio = (sorted_score[..., :base_capacity].sum(dim=-1, keepdim=True)
                 / sorted_score.sum(dim=-1, keepdim=True))
        adaptive = adaptive * ratio
    flat = adaptive.reshape(B, n * H)
    top = torch.topk(flat, k=H * base_capacity, dim=-1).indices // n
    counts = torch.zeros((B, H), dtype=top.dtype)
    counts.scatter_add_(-1, top, torch.ones_like(top))
    assert int(counts.sum()) == H * base_capacity
    cap = torch.round(counts * (1 - floor_ratio) + floor_capacity).int()
    return cap[0], sorted_idx, adaptive


def adakv_update_kv(K, Q, V, window_size=32, kernel_size=7, pooling="maxpool",
                    max_capacity_prompt=2048, floor=0.2, normalize=True):
    """``AdaKVCluster.update_kv`` (``pyramidkv_utils.py:674-757``).

    Returns (flatK [sum_h(cap_h+w), d], flatV, meta dict, pooled scores, per-head index lists).
    """
    B, Hq, L, d = Q.shape
    w = window_size
    base = max_capacity_prompt - w
    Kr, Vr = _expand(K, V, Hq)
    pooled = adakv_scores(K, Q, w, kernel_size, pooling)
    if base > pooled.size(-1):                                    # :700-703 flatten only
        meta = varlen_metadata([L] * Hq)
        return Kr.reshape(-1, d), Vr.reshape(-1, d), meta, pooled, None
    assert B == 1                                                 # :724
    cap, sorted_idx, _ = adakv_capacities(pooled, base, floor, normalize)
    ks, vs, lens, picks = [], [], [], []
    for h in range(Hq):
        sel = sorted_idx[0, h, : int(cap[h])]
        picks.append(sel)
        lens.append(int(sel.numel()) + w)
        ks.append(torch.cat([Kr[0, h, :-w][sel], Kr[0, h, -w:]], dim=0))
        vs.append(torch.cat([Vr[0, h, :-w][sel], Vr[0, h, -w:]], dim=0))
    return torch.cat(ks, 0), torch.cat(vs, 0), varlen_metadata(lens), pooled, picks


# --------------------------------------------------------------------------------------------
# a7/a8  flattened-cache append (numpy integer/byte work)
# --------------------------------------------------------------------------------------------
def update_flatten_view(cache: np.ndarray, state: np.ndarray, headlens: np.ndarray,
                        cu_headlens: np.ndarray) -> np.ndarray:
    """Rebuild the flattened varlen cache with one new row appended per head.

    Restates ``update_flatten_view_kernel`` (``csrc/csrc/cuda_api.cu:11-53``) and its host wrapper
    (``:55-85``): for head h, rows ``cache[cu[h] : cu[h]+headlens[h]]`` move to
    ``out[cu[h]+h : ...]`` and ``state[h]`` lands at row ``cu[h+1]+h``.  ``cache``/``state`` are
    16-bit patterns (fp16/bf16 viewed as uint16) or any 2-D array; only bytes move.
    """
    if headlens.dtype != np.int32 or cu_headlens.dtype != np.int32:
        raise TypeError("expected headlens / cu_headlens to be int32")          # cuda_api.cu:56-57
    total, d = cache.shape
    H = int(headlens.shape[0])
    out = np.zeros((total + H, d), dtype=cache.dtype)
    for h in range(H):
        src0 = int(cu_headlens[h])
        n = int(headlens[h])
        out[src0 + h: src0 + h + n] = cache[src0: src0 + n]
        out[int(cu_headlens[h + 1]) + h] = state[h]
    return out


# --------------------------------------------------------------------------------------------
# tie-aware comparison helpers (the comparison itself is part of the test contract)
# --------------------------------------------------------------------------------------------
def check_index_sets(pooled: torch.Tensor, idx_test: torch.Tensor, k: int, rtol: float = 0.0,
                     atol: float = 0.0):
    """Tie-aware top-k parity (SURVEY.md §7 hard parts).

    For every (b,h): each position whose oracle score exceeds the k-th largest score by more than
    ``margin = rtol*|kth| + atol`` must be selected, every selected position must score at least
    kth - margin, and the selection must have k distinct in-range entries.  Returns (ok, message).
    """
    P = pooled.float()
    B, H, n = P.shape
    idx_test = idx_test.reshape(B, H, -1).long().cpu()
    if idx_test.shape[-1] != k:
        return False, f"expected {k} indices per head, got {idx_test.shape[-1]}"
    kth = P.topk(k, dim=-1).values[..., -1:]
    for b in range(B):
        for h in range(H):
            sel = idx_test[b, h]
            if sel.min() < 0 or sel.max() >= n:
                return False, f"(b={b},h={h}) index out of range"
            if sel.unique().numel() != k:
                return False, f"(b={b},h={h}) duplicate indices"
            t = float(kth[b, h, 0])
            margin = rtol * abs(t) + atol
            must = (P[b, h] > t + margin).nonzero().flatten()
            chosen = torch.zeros(n, dtype=torch.bool)
            chosen[sel] = True
            if not bool(chosen[must].all()):
                miss = must[~chosen[must]][:5].tolist()
                return False, f"(b={b},h={h}) clear winners missing, e.g. {miss}"
            if bool((P[b, h, sel] < t - margin).any()):
                bad = sel[P[b, h, sel] < t - margin][:5].tolist()
                return False, f"(b={b},h={h}) selected below threshold, e.g. {bad}"
    return True, "ok"


# --------------------------------------------------------------------------------------------
# f1  HeadKV (AdaKV layout, externally supplied capacities)
# --------------------------------------------------------------------------------------------
def headkv_update_kv(K, Q, V, head_capacity, layer_idx, window_size=32, kernel_size=7, pooling="maxpool",
                     max_capacity_prompt=2048):
    """``HeadKVCluster.update_kv`` (``pyramidkv_utils.py:808-878``): per head keep the first
    ``head_capacity[layer_idx][h]`` positions of the descending sort of the pooled window-mean scores."""
    B, Hq, L, d = Q.shape
    w = window_size
    base = max_capacity_prompt - w
    Kr, Vr = _expand(K, V, Hq)
    pooled = adakv_scores(K, Q, w, kernel_size, pooling)
    if base > pooled.size(-1):
        return Kr.reshape(-1, d), Vr.reshape(-1, d), varlen_metadata([L] * Hq), pooled, None
    assert B == 1
    _, sorted_idx = pooled.sort(dim=-1, descending=True)
    ks, vs, lens, picks = [], [], [], []
    for h in range(Hq):
        sel = sorted_idx[0, h, : int(head_capacity[layer_idx][h])]
        picks.append(sel)
        lens.append(int(sel.numel()) + w)
        ks.append(torch.cat([Kr[0, h, :-w][sel], Kr[0, h, -w:]], dim=0))
        vs.append(torch.cat([Vr[0, h, :-w][sel], Vr[0, h, -w:]], dim=0))
    return torch.cat(ks, 0), torch.cat(vs, 0), varlen_metadata(lens), pooled, picks
