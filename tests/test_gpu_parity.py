"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, the golden vectors made
by the unmodified reference, and size-independent properties at BASELINE sizes.

Tolerances (written here because they are the contract):
  * gathers, appends, pooling, metadata: bit-exact.
  * window scores: the kernel reproduces the reference's rounding points, so the only differences
    are fp32 accumulation order inside the dot products and exp/ulp-level effects.  We require
    every score within SCORE_ULPS model-dtype ulps of the oracle's and at least SCORE_EXACT_FRAC of
    them bit-identical.
  * index sets: identical except at near-ties: every position whose oracle score beats the k-th
    value by more than TIE_RTOL (relative) must be selected and nothing below kth*(1-TIE_RTOL) may be.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from kvcache_factory_b200 import (AdaKVCluster, DynamicCacheSplitHeadFlatten, H2OKVCluster, PyramidKVCluster,
                                  SnapKVCluster, StreamingLLMKVCluster, _native)
from kvcache_factory_b200.synth import Shape, make_qkv
from oracle import kvc_oracle as O
from tests.helpers import DTYPES, from_bits, golden_cases, load_golden, score_mismatch, to_bits, ulp_of

pytestmark = pytest.mark.gpu

SCORE_ULPS = 4.0
SCORE_EXACT_FRAC = 0.90
TIE_RTOL = {torch.bfloat16: 2.0 ** -5, torch.float16: 2.0 ** -8}

IMPLS = [pytest.param(_native.IMPL_MMA_SYNC, id="mma_sync"), pytest.param(_native.IMPL_TCGEN05, id="tcgen05"),
         pytest.param(_native.IMPL_TCGEN05_FUSED, id="tcgen05_fused"),
         pytest.param(_native.IMPL_TCGEN05_HALF, id="tcgen05_half"),
         pytest.param(_native.IMPL_TCGEN05_DYN, id="tcgen05_dyn")]


def dev():
    return torch.device("cuda:0")


def _skip_if_unsupported(fn):
    try:
        return fn()
    except _native.KvcError as e:
        if e.status == -2 and "tcgen05" in str(e):
            pytest.skip(str(e))
        raise


# ------------------------------------------------------------------------------------------------
# row movers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["unrepeated", "repeated", "hf_transposed"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gather_dense_bit_exact(layout, dtype):
    shape = Shape(B=2, Hq=8, Hkv=2, L=777, d=128)
    w, n_sel = 16, 200
    g = torch.Generator().manual_seed(7)
    _, K, V = make_qkv(shape, dtype, layer_idx=3, window=w)
    idx = torch.stack([torch.randperm(shape.L - w, generator=g)[:n_sel] for _ in range(shape.B * shape.Hq)])
    idx = idx.view(shape.B, shape.Hq, n_sel)
    Kr, Vr = O.repeat_kv(K, 4), O.repeat_kv(V, 4)
    ref_k, ref_v = O.gather_and_append_window(Kr, Vr, idx, w)
    if layout == "unrepeated":
        Kd, Vd = K.to(dev()), V.to(dev())
    elif layout == "repeated":
        Kd, Vd = Kr.to(dev()), Vr.to(dev())
    else:  # memory order [B,L,H,d] viewed as [B,H,L,d], the layout RoPE output has in HF
        Kd = K.to(dev()).transpose(1, 2).contiguous().transpose(1, 2)
        Vd = V.to(dev()).transpose(1, 2).contiguous().transpose(1, 2)
        assert not Kd.is_contiguous()
    out_k, out_v = _native.gather_dense(Kd, Vd, idx.to(dev(), torch.int32).contiguous(), shape.Hq, n_sel, w)
    assert torch.equal(out_k.cpu(), ref_k) and torch.equal(out_v.cpu(), ref_v)


def test_gather_reads_value_rows_from_pinned_host_memory():
    """V is only read at the selected rows: a pinned host V gives the same bits as a device V."""
    shape = Shape(B=1, Hq=8, Hkv=2, L=1500, d=128)
    w, n_sel = 32, 300
    g = torch.Generator().manual_seed(11)
    _, K, V = make_qkv(shape, torch.bfloat16, layer_idx=1, window=w)
    idx = torch.stack([torch.randperm(shape.L - w, generator=g)[:n_sel] for _ in range(shape.Hq)]).view(1, shape.Hq, n_sel)
    ref_k, ref_v = O.gather_and_append_window(O.repeat_kv(K, 4), O.repeat_kv(V, 4), idx, w)
    out_k, out_v = _native.gather_dense(K.to(dev()), V.pin_memory(), idx.to(dev(), torch.int32).contiguous(),
                                        shape.Hq, n_sel, w)
    assert out_v.is_cuda and torch.equal(out_k.cpu(), ref_k) and torch.equal(out_v.cpu(), ref_v)
    with pytest.raises(RuntimeError, match="pinned"):
        _native.gather_dense(K.to(dev()), V, idx.to(dev(), torch.int32).contiguous(), shape.Hq, n_sel, w)


@pytest.mark.parametrize("zero_copy_v", [True, False])
def test_host_compressor_matches_device_path(zero_copy_v):
    """Host-buffer entry (pinned K/V/Q-window in, K'/V' out) == the device-resident update_kv, bit for bit."""
    from kvcache_factory_b200.host_api import HostCompressor
    shape = Shape(1, 8, 2, 2048, 128)
    w, cap, nl = 32, 256, 3
    mk = lambda l: SnapKVCluster(w, cap, 7, "avgpool")
    dev_cl, host_cl = [mk(l) for l in range(nl)], [mk(l) for l in range(nl)]
    hK, hV, hQ, ref = [], [], [], []
    for l in range(nl):
        Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=l, dist="peaked", window=w, q_window_only=True)
        Kc, Vc = dev_cl[l].update_kv(K.to(dev()), Q.to(dev()), V.to(dev()), None, 4)
        ref.append((Kc.cpu(), Vc.cpu()))
        hK.append(K.pin_memory()); hV.append(V.pin_memory()); hQ.append(Q[:, :, -w:].contiguous().pin_memory())
    for c in host_cl:
        c.async_tail = True            # as bench.py configures them: the host path must still order its copies
    hc = HostCompressor(host_cl, 4, dev(), zero_copy_v=zero_copy_v)
    hc.run(hK, hQ, hV, shape.L)
    out = hc.run(hK, hQ, hV, shape.L)   # second run: staging buffers and pinned outputs are reused
    assert all(c.async_tail for c in host_cl)
    torch.cuda.synchronize()
    for l in range(nl):
        assert torch.equal(out[l][0], ref[l][0]) and torch.equal(out[l][1], ref[l][1])
    kept = nl * shape.B * shape.Hq * cap * shape.d * 2
    full = nl * shape.B * shape.Hkv * shape.L * shape.d * 2
    qb = nl * shape.B * shape.Hq * w * shape.d * 2
    assert hc.d2h_bytes == 2 * kept
    assert hc.h2d_bytes == full + qb + (kept if zero_copy_v else full)


@pytest.mark.parametrize("name", golden_cases("streamingllm") + ["snapkv_passthrough"])
def test_streamingllm_and_passthrough_match_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    kn = c["knobs"]
    Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
    if c["method"] == "streamingllm":
        cl = StreamingLLMKVCluster(**kn)
        Kc, Vc = cl.update_kv(Kd, Qd, Vd, None, c["Hq"] // c["Hkv"])
        assert np.array_equal(to_bits(Kc), z["K_out"]) and np.array_equal(to_bits(Vc), z["V_out"])
    else:
        cl = SnapKVCluster(**kn)
        Kc, Vc = cl.update_kv(Kd, Qd, Vd, None, c["Hq"] // c["Hkv"])
        assert Kc is Kd and Vc is Vd  # same objects, as in the reference (:314-315)


def _flatten_case(H, d, lens, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    hl = np.asarray(lens, dtype=np.int32)
    cu = np.concatenate([[0], np.cumsum(hl)]).astype(np.int32)
    total = int(cu[-1])
    cache = torch.randn(total, d, generator=g).to(dtype)
    state = torch.randn(H, d, generator=g).to(dtype)
    return cache, state, hl, cu


@pytest.mark.parametrize("lens", [[5, 1, 9, 3], [0, 4, 0, 2], [1], [2080, 2011, 1999, 2100, 33, 4096, 2048, 1]])
@pytest.mark.parametrize("d", [64, 128])
def test_update_flatten_view_matches_oracles(lens, d):
    H = len(lens)
    cache, state, hl, cu = _flatten_case(H, d, lens, torch.float16)
    out = _native.update_flatten_view(cache.to(dev()), state.to(dev()), torch.from_numpy(hl).to(dev()),
                                      torch.from_numpy(cu).to(dev()))
    got = to_bits(out)
    ref = O.update_flatten_view(to_bits(cache), to_bits(state), hl, cu)
    assert got.shape == (cache.shape[0] + H, d) and np.array_equal(got, ref)
    # the plain-C restatement of cuda_api.cu agrees too
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build",
                      "libflatten_oracle.so")
    if os.path.exists(so):
        lib = ctypes.CDLL(so)
        dst = np.zeros_like(ref)
        src, st = np.ascontiguousarray(to_bits(cache)), np.ascontiguousarray(to_bits(state))
        lib.flatten_oracle_update(dst.ctypes.data_as(ctypes.c_void_p), src.ctypes.data_as(ctypes.c_void_p),
                                  st.ctypes.data_as(ctypes.c_void_p), hl.ctypes.data_as(ctypes.c_void_p),
                                  cu.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(H), ctypes.c_int32(d))
        assert np.array_equal(got, dst)


def test_update_flatten_view_errors_like_reference():
    cache, state, hl, cu = _flatten_case(2, 128, [3, 4], torch.bfloat16)
    with pytest.raises(RuntimeError, match="int32"):
        _native.update_flatten_view(cache.to(dev()), state.to(dev()), torch.from_numpy(hl).long().to(dev()),
                                    torch.from_numpy(cu).to(dev()))
    big = torch.zeros(4, 256, dtype=torch.bfloat16, device=dev())
    with pytest.raises(RuntimeError):
        _native.update_flatten_view(big, big[:2], torch.from_numpy(hl).to(dev()), torch.from_numpy(cu).to(dev()))


def test_flatten_cache_decode_steps():
    """DynamicCacheSplitHeadFlatten: prefill append then 3 decode appends with the metadata bumps of
    llama_model.py:2372-2375; every head's rows must stay in order with the new rows at the end."""
    H, d = 4, 128
    lens = [7, 3, 5, 9]
    cache, _, hl, cu = _flatten_case(H, d, lens, torch.bfloat16, seed=1)
    kc = DynamicCacheSplitHeadFlatten()
    kc.update(cache.to(dev()), cache.to(dev()).clone(), 0)
    head_lens = torch.from_numpy(hl).to(dev())
    cu_klen = torch.from_numpy(cu).to(dev())
    cu_offset = torch.arange(0, H + 1, dtype=torch.int32, device=dev())
    per_head = [cache[cu[h]:cu[h + 1]].clone() for h in range(H)]
    for step in range(3):
        new = torch.randn(1, H, 1, d).to(torch.bfloat16)
        k, v = kc.update(new.to(dev()), new.to(dev()), 0, {"head_lens": head_lens, "cu_klen": cu_klen})
        cu_klen = cu_klen + cu_offset
        head_lens = head_lens + 1
        per_head = [torch.cat([per_head[h], new[0, h]], 0) for h in range(H)]
        assert torch.equal(k.cpu(), torch.cat(per_head, 0)) and torch.equal(v.cpu(), torch.cat(per_head, 0))


@pytest.mark.parametrize("slack", [1, 5, 16])
def test_slack_flatten_cache_equals_rebuild_path(slack):
    """f3: in-place O(H*d) appends into slack-allocated head segments (re-laid out every `slack` steps) hold exactly
    the rows the reference's rebuild-every-step cache holds, and flash-attn over (segment starts, used lengths) gives
    the attention output of flash-attn over the packed cache, for 40 decode steps."""
    from flash_attn import flash_attn_varlen_func
    from flash_attn.flash_attn_interface import _flash_attn_varlen_forward
    H, d = 8, 128
    lens = [130, 7, 64, 999, 1, 33, 250, 512]
    cache, _, hl, cu = _flatten_case(H, d, lens, torch.bfloat16, seed=3)
    vcache = torch.randn(cache.shape, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    ref, slk = DynamicCacheSplitHeadFlatten(), DynamicCacheSplitHeadFlatten(slack=slack)
    hl_r, cu_r = torch.from_numpy(hl).to(dev()), torch.from_numpy(cu).to(dev())
    hl_s = hl_r.clone()
    ref.update(cache.to(dev()), vcache.to(dev()), 0)
    slk.update(cache.to(dev()), vcache.to(dev()), 0, {"head_lens": hl_s, "cu_klen": cu_r.clone()})
    off = torch.arange(0, H + 1, dtype=torch.int32, device=dev())
    cu_q = torch.arange(0, H + 1, dtype=torch.int32, device=dev())
    g = torch.Generator().manual_seed(21)
    for step in range(40):
        nk = torch.randn(1, H, 1, d, generator=g).to(torch.bfloat16).to(dev())
        nv = torch.randn(1, H, 1, d, generator=g).to(torch.bfloat16).to(dev())
        q = torch.randn(H, 1, d, generator=g).to(torch.bfloat16).to(dev())
        kr, vr = ref.update(nk, nv, 0, {"head_lens": hl_r, "cu_klen": cu_r})
        cu_r = cu_r + off
        hl_r = hl_r + 1
        ks, vs = slk.update(nk, nv, 0, {"head_lens": hl_s, "cu_klen": None})
        hl_s += 1
        cap = slk.cu_cap[0]
        assert int(cap[-1]) == ks.shape[0] and torch.equal(hl_s, hl_r)
        for h in range(H):
            a, n = int(cap[h]), int(hl_s[h])
            assert torch.equal(ks[a:a + n], kr[int(cu_r[h]):int(cu_r[h + 1])]), (step, h)
            assert torch.equal(vs[a:a + n], vr[int(cu_r[h]):int(cu_r[h + 1])]), (step, h)
        mx = int(hl_r.max())
        want = flash_attn_varlen_func(q, kr.view(-1, 1, d), vr.view(-1, 1, d), cu_q, cu_r, 1, mx, causal=True)
        got = _flash_attn_varlen_forward(q, ks.view(-1, 1, d), vs.view(-1, 1, d), cu_q, cap, 1, mx, 0.0, d ** -0.5, True,
                                         seqused_k=hl_s)[0]
        assert torch.equal(got, want), f"attention output differs at decode step {step}"


# ------------------------------------------------------------------------------------------------
# pool + top-k
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pooling,kernel", [("avgpool", 7), ("avgpool", 5), ("maxpool", 7), ("maxpool", 1), (None, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,k", [(4064, 96), (1000, 1000), (333, 1), (20000, 3932), (70000, 9000), (131040, 3964),
                                 (50000, 4096)])
def test_pool_topk_matches_torch(pooling, kernel, dtype, n, k):
    g = torch.Generator().manual_seed(n + k)
    rows = 6
    s = (torch.rand(rows, n, generator=g) ** 4 * 3).to(dtype)           # many exact ties in 16-bit
    ref_pooled = O.pool_scores(s.view(1, rows, n), kernel, pooling).view(rows, n)
    idx, pooled, sums = _native.pool_topk(s.to(dev()), kernel, pooling, k, want_pooled=True, want_sums=True)
    assert np.array_equal(to_bits(pooled), to_bits(ref_pooled)), "pooled scores are not bit-identical to ATen's"
    idx = idx.cpu().long()
    vals = ref_pooled.float().gather(1, idx)
    ref_vals = ref_pooled.float().topk(k, dim=-1).values
    assert torch.equal(vals, ref_vals), "selected values (descending) differ from torch.topk"
    for r in range(rows):
        assert idx[r].unique().numel() == k
        same = vals[r][1:] == vals[r][:-1]
        assert bool((idx[r][1:][same] > idx[r][:-1][same]).all()), "ties must be ordered by ascending position"
    assert torch.allclose(sums[:, 0].cpu(), ref_vals.sum(-1), rtol=1e-5)
    assert torch.allclose(sums[:, 1].cpu(), ref_pooled.float().sum(-1), rtol=1e-5)


def _dist(name, rows, n, g):
    if name == "ties":          # few distinct values: the threshold bin is one key with thousands of ties
        return (torch.randint(0, 12, (rows, n), generator=g).float() / 8)
    if name == "constant":      # key range 0
        return torch.full((rows, n), 0.37)
    if name == "concentrated":  # almost everything inside one adaptive bin: the one-round shortcut must fall back
        return 1.0 + torch.rand(rows, n, generator=g) * 2e-2
    if name == "zeros":         # underflowed probabilities: mostly exact zeros, a heavy tail
        x = torch.rand(rows, n, generator=g) ** 16
        return torch.where(x < 1e-3, torch.zeros(()), x)
    if name == "negative":      # L2Norm scores are negated norms
        return -(torch.rand(rows, n, generator=g) * 30 + 0.5)
    return torch.rand(rows, n, generator=g) ** 8 * 4          # "heavy": softmax-like tail


@pytest.mark.parametrize("dist", ["heavy", "ties", "constant", "concentrated", "zeros", "negative"])
@pytest.mark.parametrize("n,k", [(32736, 2016), (32768, 4096), (32736, 4097), (16385, 1), (8193, 8192), (4097, 4097),
                                 (65504, 3038), (131040, 172), (20000, 8192), (700, 128)])
def test_pool_topk_index_only_paths(dist, n, k):
    """Indices without sums: the path the clusters use (one counting round when the threshold bin fits the sort buffers,
    compaction bases from the per-CTA histograms) on distributions that stress the binning."""
    g = torch.Generator().manual_seed(n * 7 + k)
    rows = 5
    dtype = torch.bfloat16 if (n + k) % 2 == 0 else torch.float16
    s = _dist(dist, rows, n, g).to(dtype)
    for pooling, kernel in (("avgpool", 7), ("maxpool", 5), (None, 1)):
        ref_pooled = O.pool_scores(s.view(1, rows, n), kernel, pooling).view(rows, n)
        idx, _, _ = _native.pool_topk(s.to(dev()), kernel, pooling, k)
        idx = idx.cpu().long()
        assert int(idx.min()) >= 0 and int(idx.max()) < n
        vals = ref_pooled.float().gather(1, idx)
        assert torch.equal(vals, ref_pooled.float().topk(k, dim=-1).values), (dist, pooling, "values differ from torch.topk")
        for r in range(rows):
            assert idx[r].unique().numel() == k
            same = vals[r][1:] == vals[r][:-1]
            assert bool((idx[r][1:][same] > idx[r][:-1][same]).all()), "ties must be ordered by ascending position"


def test_pool_topk_per_row_k():
    g = torch.Generator().manual_seed(5)
    rows, n = 8, 5000
    s = torch.rand(rows, n, generator=g).to(torch.bfloat16)
    kk = torch.tensor([0, 1, 17, 4999, 5000, 2048, 100, 3], dtype=torch.int32)
    idx, _, _ = _native.pool_topk(s.to(dev()), 7, "maxpool", 5000, k_per_row=kk.to(dev()))
    ref_pooled = O.pool_scores(s.view(1, rows, n), 7, "maxpool").view(rows, n).float()
    for r in range(rows):
        k = int(kk[r])
        if k:
            assert torch.equal(ref_pooled[r][idx[r, :k].cpu().long()], ref_pooled[r].topk(k).values)


def test_pool_topk_errors():
    s = torch.zeros(2, 100, dtype=torch.bfloat16, device=dev())
    with pytest.raises(ValueError, match="Pooling method not supported"):
        _native.pool_topk(s, 5, "medianpool", 10)
    with pytest.raises(_native.KvcError):
        _native.pool_topk(s, 4, "avgpool", 10)        # even kernels change the length in the reference
    with pytest.raises(_native.KvcError):
        _native.pool_topk(s, 5, "avgpool", 101)       # k > n: torch.topk raises too


# ------------------------------------------------------------------------------------------------
# window scores
# ------------------------------------------------------------------------------------------------
SCORE_CASES = [
    # B, Hq, Hkv, L, d, w
    (1, 8, 2, 512, 128, 32),
    (2, 4, 4, 300, 128, 8),       # repeated layout (g = 1), ragged L
    (1, 4, 1, 641, 128, 32),
    (1, 16, 2, 1500, 128, 32),    # g = 8 (70B geometry): 256 stacked rows
    (1, 6, 3, 333, 64, 16),       # d = 64
    (1, 4, 2, 200, 128, 20),      # window not a multiple of 16
    (1, 32, 8, 4096, 128, 32),    # cfg-1 geometry
    (1, 32, 8, 2048, 128, 8),     # LongBench driver default window 8: 32 of 128 accumulator rows are real
    (1, 32, 8, 2048, 128, 16),    # 64 of 128 rows
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("case", SCORE_CASES, ids=lambda c: "B%d_Hq%d_Hkv%d_L%d_d%d_w%d" % c)
def test_window_scores_match_oracle(impl, dtype, reduce, case):
    B, Hq, Hkv, L, d, w = case
    shape = Shape(B, Hq, Hkv, L, d)
    Q, K, V = make_qkv(shape, dtype, layer_idx=1, dist="peaked", window=w)
    ref = O.window_scores(O.repeat_kv(K, Hq // Hkv), Q, w, reduce)
    got = _skip_if_unsupported(lambda: _native.window_scores(Q.to(dev()), K.to(dev()), w, reduce, impl))
    torch.cuda.synchronize()
    assert got.shape == ref.shape and got.dtype == dtype
    bad, worst = score_mismatch(got, ref, dtype, SCORE_ULPS)
    exact = float((got.cpu() == ref).float().mean())
    assert bad == 0.0, f"{bad:.2%} of scores off by more than {SCORE_ULPS} ulps (worst {worst:.1f}x)"
    assert exact >= SCORE_EXACT_FRAC, f"only {exact:.2%} of scores bit-identical to the oracle"


@pytest.mark.parametrize("impl", IMPLS)
def test_window_scores_hf_query_layout(impl):
    """query_states arrives as a transposed view of [B,L,Hq,d]: same result."""
    shape = Shape(1, 8, 2, 700, 128)
    Q, K, _ = make_qkv(shape, torch.bfloat16, layer_idx=2, dist="peaked", window=32)
    a = _skip_if_unsupported(lambda: _native.window_scores(Q.to(dev()), K.to(dev()), 32, "sum", impl))
    Qt = Q.to(dev()).transpose(1, 2).contiguous().transpose(1, 2)
    Kt = K.to(dev()).transpose(1, 2).contiguous().transpose(1, 2)
    b = _native.window_scores(Qt, Kt, 32, "sum", impl)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# H2O (all-query) scores
# ------------------------------------------------------------------------------------------------
H2O_CASES = [
    # B, Hq, Hkv, L, w
    (1, 4, 2, 384, 32),
    (2, 4, 1, 1000, 16),     # ragged L, g = 4, window inside the last tile
    (1, 2, 2, 129, 8),       # one key past a tile boundary
    (1, 8, 2, 2048, 64),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", H2O_CASES, ids=lambda c: "B%d_Hq%d_Hkv%d_L%d_w%d" % c)
def test_h2o_scores_match_oracle(dtype, case):
    B, Hq, Hkv, L, w = case
    shape = Shape(B, Hq, Hkv, L, 128)
    Q, K, V = make_qkv(shape, dtype, layer_idx=6, dist="peaked", window=w)
    ref = O.window_scores(O.repeat_kv(K, Hq // Hkv), Q, w, "sum", all_queries=True)
    got = _native.h2o_scores(Q.to(dev()), K.to(dev()), w)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and got.dtype == dtype
    bad, worst = score_mismatch(got, ref, dtype, SCORE_ULPS)
    exact = float((got.cpu() == ref).float().mean())
    assert bad == 0.0, f"{bad:.2%} of H2O scores off by more than {SCORE_ULPS} ulps (worst {worst:.1f}x)"
    assert exact >= 0.80, f"only {exact:.2%} of H2O scores bit-identical to the oracle"


@pytest.mark.parametrize("name", golden_cases("h2o"))
def test_h2o_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    _check_dense(H2OKVCluster(**c["knobs"]), c, z, Q, K, V, _native.IMPL_AUTO)


# ------------------------------------------------------------------------------------------------
# whole update_kv against the reference's vectors
# ------------------------------------------------------------------------------------------------
def _check_dense(cluster, c, z, Q, K, V, impl):
    dtype = DTYPES[c["dtype"]]
    g = c["Hq"] // c["Hkv"]
    cluster.keep_debug = True
    cluster.score_impl = impl
    Kc, Vc = _skip_if_unsupported(lambda: cluster.update_kv(K.to(dev()), Q.to(dev()), V.to(dev()), None, g))
    torch.cuda.synchronize()
    assert tuple(Kc.shape) == tuple(z["K_out"].shape) and Kc.dtype == dtype
    idx = cluster.debug["indices"].cpu().long()
    pooled_ref = from_bits(z["pooled"], dtype)
    k = z["idx"].shape[-1]
    ok, msg = O.check_index_sets(pooled_ref, idx, k, rtol=TIE_RTOL[dtype])
    assert ok, msg
    # given the indices, the gathered rows are bit-exact
    w = c["knobs"]["window_size"]
    Kr, Vr = O.repeat_kv(K, g), O.repeat_kv(V, g)
    gk, gv = O.gather_and_append_window(Kr, Vr, idx, w)
    assert torch.equal(Kc.cpu(), gk) and torch.equal(Vc.cpu(), gv)
    # how close the selection is to the reference's own (reported, ties make 100% unattainable)
    ref_sets = z["idx"]
    agree = np.mean([len(set(idx[b, h].tolist()) & set(ref_sets[b, h].tolist())) / k
                     for b in range(idx.shape[0]) for h in range(idx.shape[1])])
    assert agree > 0.9, f"only {agree:.1%} of the reference's picks were selected"


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", golden_cases("snapkv"))
def test_snapkv_matches_reference_vectors(name, impl):
    if name == "snapkv_passthrough":
        pytest.skip("covered by the pass-through test")
    c, z, (Q, K, V) = load_golden(name)
    _check_dense(SnapKVCluster(**c["knobs"]), c, z, Q, K, V, impl)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", golden_cases("pyramidkv"))
def test_pyramidkv_matches_reference_vectors(name, impl):
    c, z, (Q, K, V) = load_golden(name)
    _check_dense(PyramidKVCluster(**c["knobs"]), c, z, Q, K, V, impl)


@pytest.mark.parametrize("method", ["snapkv", "pyramidkv", "streamingllm"])
@pytest.mark.parametrize("layout", ["full_q", "window_q", "hf_transposed"])
def test_fused_one_call_path_equals_three_calls(method, layout):
    """kvc_compress_dense (the synchronous default: one library call per update_kv) launches the same three kernels as
    the separate window_scores / pool_topk / gather_dense calls: identical bits, also from the HF tensor layouts."""
    shape = Shape(2, 8, 2, 1777, 128)
    w, cap = 32, 256
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=5, dist="peaked", window=w)
    Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
    if layout == "hf_transposed":
        Qd = Qd.transpose(1, 2).contiguous().transpose(1, 2)
        Kd = Kd.transpose(1, 2).contiguous().transpose(1, 2)
        Vd = Vd.transpose(1, 2).contiguous().transpose(1, 2)
    mk = {"snapkv": lambda: SnapKVCluster(w, cap, 7, "avgpool"),
          "pyramidkv": lambda: PyramidKVCluster(num_hidden_layers=8, window_size=w, max_capacity_prompt=cap, kernel_size=5,
                                                pooling="maxpool", layer_idx=3),
          "streamingllm": lambda: StreamingLLMKVCluster(w, cap)}[method]
    a, b = mk(), mk()
    b.fused_call = False
    if layout == "window_q":
        if method == "streamingllm":
            pytest.skip("StreamingLLM reads no queries")
        Ka, Va = a.update_kv_window(Kd, Qd[:, :, -w:].contiguous(), Vd, 4, shape.L)
        Kb, Vb = b.update_kv_window(Kd, Qd[:, :, -w:].contiguous(), Vd, 4, shape.L)
    else:
        Ka, Va = a.update_kv(Kd, Qd, Vd, None, 4)
        Kb, Vb = b.update_kv(Kd, Qd, Vd, None, 4)
    torch.cuda.synchronize()
    assert Ka.shape == Kb.shape and torch.equal(Ka, Kb) and torch.equal(Va, Vb)
    with pytest.raises(ValueError, match="Pooling method not supported"):
        SnapKVCluster(w, cap, 7, "medianpool").update_kv(Kd, Qd, Vd, None, 4)


@pytest.mark.parametrize("pooling,floor,normalize", [("maxpool", 0.2, True), ("avgpool", 0.5, False)])
@pytest.mark.parametrize("hf_layout", [False, True])
def test_adakv_one_call_equals_separate_calls(pooling, floor, normalize, hf_layout):
    """kvc_adakv_prefill (the default: one library call per AdaKV update_kv) against the five separate calls:
    identical capacities, metadata and flattened rows."""
    shape = Shape(1, 8, 2, 3001, 128)
    w = 32
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=2, dist="peaked", window=w)
    Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
    if hf_layout:
        Qd = Qd.transpose(1, 2).contiguous().transpose(1, 2)
        Kd = Kd.transpose(1, 2).contiguous().transpose(1, 2)
        Vd = Vd.transpose(1, 2).contiguous().transpose(1, 2)
    mk = lambda: AdaKVCluster(window_size=w, kernel_size=7, pooling=pooling, max_capacity_prompt=256 + w, floor=floor,
                              normalize=normalize, layer_idx=0, num_hidden_layers=4)
    a, b = mk(), mk()
    a.keep_debug = b.keep_debug = True
    b.fused_call = False
    Ka, Va = a.update_kv(Kd, Qd, Vd)
    Kb, Vb = b.update_kv(Kd, Qd, Vd)
    torch.cuda.synchronize()
    assert Ka.shape == Kb.shape and Ka.is_contiguous() and torch.equal(Ka, Kb) and torch.equal(Va, Vb)
    for attr in ("head_lens", "cu_klen", "cu_headlens"):
        assert torch.equal(getattr(a, attr), getattr(b, attr)), attr
    assert (a.klen_sum, a.max_seqlen_k) == (b.klen_sum, b.max_seqlen_k)
    assert torch.equal(a.debug["pooled"].view(-1), b.debug["pooled"].view(-1))
    caps = (a.head_lens - w).tolist()
    for h, c in enumerate(caps):      # entries past a head's capacity are unspecified
        assert torch.equal(a.debug["indices"][h, :c], b.debug["indices"][h, :c]), h


def test_repeated_layout_gives_same_result_as_unrepeated():
    shape = Shape(1, 8, 2, 600, 128)
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=4, dist="peaked", window=32)
    a = SnapKVCluster(32, 128, 7, "avgpool")
    b = SnapKVCluster(32, 128, 7, "avgpool")
    a.score_impl = b.score_impl = _native.IMPL_MMA_SYNC
    Ka, Va = a.update_kv(K.to(dev()), Q.to(dev()), V.to(dev()), None, 4)
    Kb, Vb = b.update_kv(O.repeat_kv(K, 4).to(dev()), Q.to(dev()), O.repeat_kv(V, 4).to(dev()), None, 4)
    assert torch.equal(Ka, Kb) and torch.equal(Va, Vb)


def test_cluster_errors_like_reference():
    with pytest.raises(AssertionError):
        SnapKVCluster(window_size=64, max_capacity_prompt=64)
    shape = Shape(1, 4, 2, 256, 128)
    Q, K, V = make_qkv(shape, torch.bfloat16, window=16)
    cl = SnapKVCluster(16, 64, 5, "medianpool")
    with pytest.raises(ValueError, match="Pooling method not supported"):
        cl.update_kv(K.to(dev()), Q.to(dev()), V.to(dev()), None, 2)
    with pytest.raises(AssertionError):
        SnapKVCluster(16, 64, 5, "avgpool").update_kv(K.to(dev())[:, :, :100], Q.to(dev()), V.to(dev()), None, 2)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        SnapKVCluster(16, 64, 5, "avgpool").update_kv(K, Q, V, None, 2)
    with pytest.raises(TypeError):
        SnapKVCluster(16, 64, 5, "avgpool").update_kv(K.float().to(dev()), Q.float().to(dev()), V.float().to(dev()),
                                                      None, 2)


# ------------------------------------------------------------------------------------------------
# AdaKV
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [pytest.param(_native.IMPL_AUTO, id="auto")] + IMPLS)
@pytest.mark.parametrize("name", golden_cases("adakv"))
def test_adakv_matches_reference_vectors(name, impl):
    c, z, (Q, K, V) = load_golden(name)
    kn = c["knobs"]
    dtype = DTYPES[c["dtype"]]
    w, H, d = kn["window_size"], c["Hq"], c["d"]
    cl = AdaKVCluster(**kn)
    cl.keep_debug = True
    cl.score_impl = impl
    fk, fv = _skip_if_unsupported(lambda: cl.update_kv(K.to(dev()), Q.to(dev()), V.to(dev())))
    torch.cuda.synchronize()
    # metadata contract (dtypes, shapes, internal consistency)
    for a in ("head_lens", "cu_klen", "cu_headlens", "cu_qlen", "cu_offset", "cu_head_offset", "layer_qlens"):
        t = getattr(cl, a)
        assert t.dtype == torch.int32 and tuple(t.shape) == tuple(z["meta_" + a].shape), a
    for a in ("cu_qlen", "cu_offset", "cu_head_offset", "layer_qlens"):
        assert np.array_equal(getattr(cl, a).cpu().numpy(), z["meta_" + a]), a
    hl = cl.head_lens.cpu().numpy()
    cu = cl.cu_klen.cpu().numpy()
    assert np.array_equal(cu, np.concatenate([[0], np.cumsum(hl)])) and cl.klen_sum == int(hl.sum()) == fk.shape[0]
    assert np.array_equal(cl.cu_headlens.cpu().numpy(), np.cumsum(hl)) and cl.max_seqlen_k == int(hl.max())
    assert cl.qlen_sum == H
    if name == "adakv_nocompress":
        assert np.array_equal(hl, z["meta_head_lens"])
        assert np.array_equal(to_bits(fk), z["K_out"]) and np.array_equal(to_bits(fv), z["V_out"])
        return
    # capacities: torch's global topk breaks ties arbitrarily, so a head may differ from the reference by
    # at most the number of its scores tied at the global threshold; the total is tie-independent up to rounding
    ref_hl = z["meta_head_lens"]
    pooled_ref = from_bits(z["pooled"], dtype)
    cap_ref, _, adaptive = O.adakv_capacities(pooled_ref, kn["max_capacity_prompt"] - w, kn["floor"], kn["normalize"])
    flat = adaptive.float().reshape(-1)
    kth = flat.topk(H * (kn["max_capacity_prompt"] - w)).values[-1]
    ties = (adaptive.float()[0] == kth).sum(-1).numpy()
    assert np.all(np.abs(hl - ref_hl) <= ties * (1 - kn["floor"]) + 1), (hl, ref_hl, ties)
    assert abs(int(hl.sum()) - int(ref_hl.sum())) <= H
    # per head: the kept rows are that head's top-cap rows (tie-aware) followed by the window, bit-exact
    g = c["Hq"] // c["Hkv"]
    Kr, Vr = O.repeat_kv(K, g), O.repeat_kv(V, g)
    idx = cl.debug["indices"].cpu().long()
    for h in range(H):
        cap = int(hl[h]) - w
        sel = idx[h, :cap]
        ok, msg = O.check_index_sets(pooled_ref[:, h:h + 1], sel.view(1, 1, -1), cap, rtol=TIE_RTOL[dtype])
        assert ok, f"head {h}: {msg}"
        rows_k = torch.cat([Kr[0, h, :-w][sel], Kr[0, h, -w:]], 0)
        rows_v = torch.cat([Vr[0, h, :-w][sel], Vr[0, h, -w:]], 0)
        assert torch.equal(fk[cu[h]:cu[h + 1]].cpu(), rows_k) and torch.equal(fv[cu[h]:cu[h + 1]].cpu(), rows_v)


@pytest.mark.parametrize("impl", [pytest.param(_native.IMPL_AUTO, id="auto")] + IMPLS)
@pytest.mark.parametrize("name", golden_cases("headkv"))
def test_headkv_matches_reference_vectors(name, impl):
    from kvcache_factory_b200 import HeadKVCluster
    c, z, (Q, K, V) = load_golden(name)
    kn = dict(c["knobs"])
    dtype = DTYPES[c["dtype"]]
    w, H = kn["window_size"], c["Hq"]
    cl = HeadKVCluster(**kn)
    cl.keep_debug = True
    cl.score_impl = impl
    fk, fv = _skip_if_unsupported(lambda: cl.update_kv(K.to(dev()), Q.to(dev()), V.to(dev())))
    torch.cuda.synchronize()
    # capacities are inputs here, so the metadata is exact
    for a in ("head_lens", "cu_klen", "cu_headlens", "cu_qlen", "cu_offset", "cu_head_offset", "layer_qlens"):
        assert np.array_equal(getattr(cl, a).cpu().numpy(), z["meta_" + a]), a
    assert [cl.klen_sum, cl.qlen_sum, cl.max_seqlen_k] == z["meta_scalars"].tolist()
    pooled_ref = from_bits(z["pooled"], dtype)
    g = c["Hq"] // c["Hkv"]
    Kr, Vr = O.repeat_kv(K, g), O.repeat_kv(V, g)
    idx = cl.debug["indices"].cpu().long()
    cu = cl.cu_klen.cpu().numpy()
    for h in range(H):
        cap = int(cl.head_lens[h]) - w
        sel = idx[h, :cap]
        ok, msg = O.check_index_sets(pooled_ref[:, h:h + 1], sel.view(1, 1, -1), cap, rtol=TIE_RTOL[dtype])
        assert ok, f"head {h}: {msg}"
        assert torch.equal(fk[cu[h]:cu[h + 1]].cpu(), torch.cat([Kr[0, h, :-w][sel], Kr[0, h, -w:]], 0))
        assert torch.equal(fv[cu[h]:cu[h + 1]].cpu(), torch.cat([Vr[0, h, :-w][sel], Vr[0, h, -w:]], 0))


# ------------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
def test_pyramidkv_full_size_properties(impl):
    """cfg-2 geometry (Llama-3-8B heads, L=32768, cap 2048, w=32, avgpool 7), two layers."""
    shape = Shape(1, 32, 8, 32768, 128)
    w, cap = 32, 2048
    for layer in (0, 31):
        Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=layer, dist="peaked", window=w, q_window_only=True)
        Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
        cl = PyramidKVCluster(num_hidden_layers=32, window_size=w, max_capacity_prompt=cap, kernel_size=7,
                              pooling="avgpool", layer_idx=layer)
        cl.keep_debug = True
        cl.score_impl = impl
        Kc, Vc = _skip_if_unsupported(lambda: cl.update_kv(Kd, Qd, Vd, None, 4))
        torch.cuda.synchronize()
        budget = O.pyramidkv_budget(shape.L, w, cap, 32, layer)
        assert Kc.shape == (1, 32, budget + w, 128)
        idx = cl.debug["indices"].long()
        scores = cl.debug["scores"].float()
        # probabilities of each window row sum to <= 1 over the non-window keys
        assert float(scores.sum(-1).max()) <= w * (1 + 2.0 ** -6)
        assert float(scores.min()) >= 0.0
        # indices: distinct, in range, ordered by descending pooled score
        assert int(idx.min()) >= 0 and int(idx.max()) < shape.L - w
        srt = idx.sort(-1).values
        assert bool((srt[..., 1:] != srt[..., :-1]).all())
        pooled = cl.debug["pooled"].float()
        pv = pooled.gather(-1, idx)
        assert bool((pv[..., 1:] <= pv[..., :-1]).all())
        kth = pv[..., -1:]
        assert bool(((pooled > kth).sum(-1) <= budget).all()), "something above the threshold was left out"
        # gather round trip: output rows are exactly the addressed input rows (checked with torch on the GPU)
        Kr = Kd.repeat_interleave(4, dim=1)
        assert torch.equal(Kc[:, :, :budget], Kr.gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, 128)))
        assert torch.equal(Kc[:, :, budget:], Kr[:, :, -w:])
        Vr = Vd.repeat_interleave(4, dim=1)
        assert torch.equal(Vc[:, :, :budget], Vr.gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, 128)))
        # the planted heavy hitters (1 % of positions, +4*qbar) dominate the selection
        del Kr, Vr


@pytest.mark.parametrize("async_scores", [False, True])
def test_async_tail_matches_synchronous_path(async_scores):
    """async_tail: select and gather of each layer run on rotating side streams (selects of consecutive
    layers overlap); after wait_for_compression() the results are bit-identical to the synchronous path."""
    from kvcache_factory_b200 import wait_for_compression
    shape = Shape(1, 8, 2, 20000, 128)         # n = 19968 scores per row: the 512-thread cluster select
    w, cap, nl = 32, 2048, 7
    ins, ref = [], []
    for l in range(nl):
        Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=l, dist="peaked", window=w, q_window_only=True)
        Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
        ins.append((Kd, Qd, Vd))
        cl = PyramidKVCluster(num_hidden_layers=nl, window_size=w, max_capacity_prompt=cap, kernel_size=7,
                              pooling="avgpool", layer_idx=l)
        if async_scores:
            cl.score_impl = _native.IMPL_TCGEN05_HALF   # same kernel as the asynchronous path: bit-comparable
        ref.append(cl.update_kv(Kd, Qd, Vd, None, 4))
    torch.cuda.synchronize()
    outs = []
    for rep in range(2):                          # second round reuses the streams
        outs = []
        for l in range(nl):
            cl = PyramidKVCluster(num_hidden_layers=nl, window_size=w, max_capacity_prompt=cap, kernel_size=7,
                                  pooling="avgpool", layer_idx=l)
            cl.async_tail = True
            cl.async_scores = async_scores     # scoring too leaves the caller's stream (half-footprint kernel, 2 streams)
            outs.append(cl.update_kv(*ins[l], None, 4))
        wait_for_compression(dev())
        torch.cuda.current_stream().synchronize()
        for l in range(nl):
            assert torch.equal(outs[l][0], ref[l][0]) and torch.equal(outs[l][1], ref[l][1]), f"layer {l} round {rep}"


def test_scoring_implementations_agree_on_random_shapes():
    """Every tcgen05 scoring kernel against the mma.sync kernel on ragged random shapes (chunk boundaries,
    partial tiles, padded row blocks, small windows): same rounding points, so at most 1 ulp apart and
    almost always identical."""
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(24):
        B = int(rng.integers(1, 3)); Hkv = int(rng.choice([1, 2, 4])); g = int(rng.choice([1, 2, 4, 8]))
        w = int(rng.choice([8, 16, 32])); L = int(rng.integers(w + 130, 3000))
        dtype = torch.bfloat16 if trial % 3 else torch.float16
        shape = Shape(B, Hkv * g, Hkv, L, 128)
        Q, K, _ = make_qkv(shape, dtype, layer_idx=trial, dist="peaked", window=w, q_window_only=True)
        Qd, Kd = Q.to(dev()), K.to(dev())
        ref = _native.window_scores(Qd, Kd, w, "sum", _native.IMPL_MMA_SYNC, seq_len=L)
        for impl in (_native.IMPL_TCGEN05, _native.IMPL_TCGEN05_FUSED, _native.IMPL_TCGEN05_HALF, _native.IMPL_TCGEN05_DYN):
            try:
                out = _native.window_scores(Qd, Kd, w, "sum", impl, seq_len=L)
            except _native.KvcError as e:
                if e.status == -2:
                    continue
                raise
            a, b = to_bits(ref.cpu()).astype(np.int64), to_bits(out.cpu()).astype(np.int64)
            assert np.abs(a - b).max() <= 1, (trial, shape, w, impl)
            assert (a == b).mean() > 0.99, (trial, shape, w, impl)
            checked += 1
    assert checked >= 55


# ------------------------------------------------------------------------------------------------
# f4: L2Norm (scores from the key norms; select + gather reused)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases("l2norm"))
def test_l2norm_matches_reference_vectors(name):
    from kvcache_factory_b200 import L2NormCluster
    c, z, (Q, K, V) = load_golden(name)
    dtype = DTYPES[c["dtype"]]
    g = c["Hq"] // c["Hkv"]
    cl = L2NormCluster(**c["knobs"])
    cl.keep_debug = True
    Kd, Qd, Vd = K.to(dev()), Q.to(dev()), V.to(dev())
    Kc, Vc = cl.update_kv(Kd, Qd, Vd, None, g)
    torch.cuda.synchronize()
    if name == "l2norm_skip":
        assert Kc is Kd and Vc is Vd                     # same objects on the skip-layer branch (:416-417)
        return
    cap = c["knobs"]["max_capacity_prompt"]
    assert tuple(Kc.shape) == tuple(z["K_out"].shape) and Kc.dtype == dtype
    # norms: fp32 sum of squares + sqrt + one rounding; the accumulation order differs from ATen's -> <= 1 ulp
    norms_ref = from_bits(z["norms"], dtype)
    got = -cl.debug["scores"].cpu()
    assert float(((got.float() - norms_ref.float()).abs() / norms_ref.float()).max()) <= 2 * ulp_of(dtype)
    assert float((got == norms_ref).float().mean()) > 0.98
    # the cap smallest norms (tie-aware, on the reference's own norms), ascending; gathered rows bit-exact
    idx = cl.debug["indices"].cpu().long()
    ok, msg = O.check_index_sets(-norms_ref, idx, cap, rtol=2 * ulp_of(dtype))
    assert ok, msg
    sel = norms_ref.float().gather(-1, idx)
    assert bool((sel[..., 1:] >= sel[..., :-1] - 2 * ulp_of(dtype) * sel[..., 1:]).all()), "not in ascending-norm order"
    Kr, Vr = O.repeat_kv(K, g), O.repeat_kv(V, g)
    gi = idx.unsqueeze(-1).expand(-1, -1, -1, c["d"])
    assert torch.equal(Kc.cpu(), Kr.gather(2, gi)) and torch.equal(Vc.cpu(), Vr.gather(2, gi))


# ------------------------------------------------------------------------------------------------
# f4: window scores from supplied statistics (log-sum-exp of the window rows), pass 2 only
# ------------------------------------------------------------------------------------------------
LSE_CASES = [(1, 8, 2, 777, 32), (2, 32, 8, 400, 8), (1, 16, 2, 1500, 32), (1, 32, 8, 4096, 32), (1, 32, 8, 2048, 16)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", LSE_CASES, ids=lambda c: "B%d_Hq%d_Hkv%d_L%d_w%d" % c)
def test_window_scores_from_exact_lse_match_two_pass_kernel(dtype, case):
    """With the log-sum-exp of the reference's OWN (dtype-rounded) logits the pass-2-only kernel must reproduce the
    two-pass scores: c_row = -lse*log2e instead of -M*log2e - log2 Z, i.e. the same number up to fp32 rounding."""
    B, Hq, Hkv, L, w = case
    shape = Shape(B, Hq, Hkv, L, 128)
    Q, K, _ = make_qkv(shape, dtype, layer_idx=2, dist="peaked", window=w)
    Kr = O.repeat_kv(K, Hq // Hkv)
    logits = torch.matmul(Q[..., -w:, :], Kr.transpose(2, 3)) / (128 ** 0.5)
    logits[:, :, -w:, -w:] += O._block_causal_bias(w, logits.dtype)[None, None]
    lse = torch.logsumexp(logits.float(), dim=-1)                      # [B,Hq,w]
    ref = O.window_scores(Kr, Q, w, "sum")
    got = _native.window_scores(Q.to(dev()), K.to(dev()), w, "sum", lse=lse.to(dev()))
    torch.cuda.synchronize()
    bad, worst = score_mismatch(got, ref, dtype, SCORE_ULPS)
    exact = float((got.cpu() == ref).float().mean())
    assert bad == 0.0, f"{bad:.2%} of scores off by more than {SCORE_ULPS} ulps (worst {worst:.1f}x)"
    assert exact >= SCORE_EXACT_FRAC, f"only {exact:.2%} of scores bit-identical to the oracle"
    # window-only queries take the same path
    got_w = _native.window_scores(Q[:, :, -w:].contiguous().to(dev()), K.to(dev()), w, "sum", seq_len=L, lse=lse.to(dev()))
    assert torch.equal(got, got_w)


def test_window_scores_from_flash_attention_lse():
    """The intended source: softmax_lse of the prefill's flash attention.  Its logits are NOT rounded to the model dtype,
    so this is an approximation of the reference (stated tolerance: scores within 2 % relative or 1e-4 absolute, top-k
    sets >= 97 % identical), not parity."""
    flash_attn = pytest.importorskip("flash_attn")
    shape = Shape(1, 32, 8, 8192, 128)
    w, cap = 32, 1024
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=8, dist="peaked", window=w)
    Qd, Kd, Vd = Q.to(dev()), K.to(dev()), V.to(dev())
    _, lse, _ = flash_attn.flash_attn_func(Qd.transpose(1, 2), Kd.transpose(1, 2), Vd.transpose(1, 2), causal=True,
                                           return_attn_probs=True)
    wl = lse[:, :, -w:].float().contiguous()
    a = SnapKVCluster(w, cap, 7, "avgpool"); a.keep_debug = True
    b = SnapKVCluster(w, cap, 7, "avgpool"); b.keep_debug = True
    Ka, Va = a.update_kv(Kd, Qd, Vd, None, 4)
    Kb, Vb = b.update_kv(Kd, Qd, Vd, None, 4, window_lse=wl)
    torch.cuda.synchronize()
    sa, sb = a.debug["scores"].float(), b.debug["scores"].float()
    assert bool(((sa - sb).abs() <= 0.02 * sa.abs() + 1e-4).all())
    ia, ib = a.debug["indices"].cpu(), b.debug["indices"].cpu()
    agree = np.mean([len(set(ia[0, h].tolist()) & set(ib[0, h].tolist())) / (cap - w) for h in range(32)])
    assert agree >= 0.97, agree
    assert Kb.shape == Ka.shape and torch.equal(Kb[:, :, -w:], Ka[:, :, -w:])


# ------------------------------------------------------------------------------------------------
# f4: ThinK (SnapKV compression + query-driven key-channel pruning of the compressed keys)
# ------------------------------------------------------------------------------------------------
def _check_think(cl, Kd, Qd, Vd, g, d, dtype):
    res = cl.update_think(Kd, Qd, Vd, None, g)
    Kc, Vc = cl.update_kv(Kd, Qd, Vd, None, g)                      # the compressed cache the pruner saw (deterministic)
    torch.cuda.synchronize()
    pruned, recent, mask, Vt = res
    rows, rs, k = Kc.shape[2], min(cl.recent_size, Kc.shape[2]), int(d * cl.ratio)
    assert torch.equal(Vt, Vc) and torch.equal(recent, Kc[:, :, rows - rs:, :])
    assert mask.dtype == torch.bool and tuple(mask.shape) == (Kc.shape[0], Kc.shape[1], d)
    assert bool((mask.sum(-1) == d - k).all())
    # the reference pruner on the same compressed keys (CPU restatement): channel scores differ from the kernel's only in
    # the fp32 summation order -> every channel clearly below / above the k-th smallest score is dropped / kept
    _, _, mask_ref, key = O.think_key_pruner(Kc.cpu(), Qd.cpu(), cl.recent_size, cl.ratio)
    keyf = key.float()
    kth = keyf.sort(-1).values[..., k - 1:k] if k > 0 else None
    m = mask.cpu()
    if k > 0:
        tol = 4 * ulp_of(dtype)
        assert not bool((m & (keyf < kth * (1 - tol))).any()), "a channel clearly below the threshold was kept"
        assert not bool((~m & (keyf > kth * (1 + tol))).any()), "a channel clearly above the threshold was dropped"
        assert float((m == mask_ref).float().mean()) > 0.97
    # given the mask: kept channels in channel order, bit for bit (:26)
    mk = m.unsqueeze(2).expand(-1, -1, rows - rs, -1)
    want = Kc.cpu()[:, :, :rows - rs, :][mk].reshape(1, -1, rows - rs, d - k)
    assert tuple(pruned.shape) == tuple(want.shape) and torch.equal(pruned.cpu(), want)
    return mask, Kc


@pytest.mark.parametrize("name", golden_cases("think"))
def test_think_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    dtype = DTYPES[c["dtype"]]
    g = c["Hq"] // c["Hkv"]
    cl = SnapKVCluster(**c["knobs"])
    Kd, Qd, Vd = K.to(dev()), Q.to(dev()), V.to(dev())
    if name == "think_skip":
        res = cl.update_think(Kd, Qd, Vd, None, g)
        assert len(res) == 2 and res[0] is Kd and res[1] is Vd        # pass-through below the capacity (:357-358)
        return
    mask, Kc = _check_think(cl, Kd, Qd, Vd, g, c["d"], dtype)
    if np.array_equal(to_bits(Kc.cpu()), z["K_out"]):                 # same compressed keys as the reference's run
        key = from_bits(z["think_key"], dtype).float()
        k = int(c["d"] * c["knobs"]["ratio"])
        kth = key.sort(-1).values[..., k - 1:k]
        m, mref = mask.cpu(), torch.from_numpy(z["think_mask"])
        near = (key - kth).abs() <= 4 * ulp_of(dtype) * kth.abs()
        assert bool(((m == mref) | near).all())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_think_larger_shape_and_hf_layout(dtype):
    shape = Shape(2, 8, 2, 2500, 128)
    Q, K, V = make_qkv(shape, dtype, layer_idx=1, dist="peaked", window=32)
    Qd = Q.to(dev()).transpose(1, 2).contiguous().transpose(1, 2)     # HF hands a transposed view
    cl = SnapKVCluster(32, 512, 7, "maxpool", recent_size=48, ratio=0.3)
    _check_think(cl, K.to(dev()), Qd, V.to(dev()), 4, 128, dtype)
