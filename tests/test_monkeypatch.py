"""replace_llama / replace_mistral (reference pyramidkv/monkeypatch.py:19-145) under the installed transformers.

CPU tests drive the hook logic with an ORACLE-backed cluster (the product clusters run on CUDA only);
the GPU test runs the real CUDA path inside a tiny random Llama and checks the cached rows against
the oracle applied to the tensors the hook handed to ``update_kv``.
"""
import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")
from transformers import LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM  # noqa: E402

import kvcache_factory_b200 as kvc  # noqa: E402
from kvcache_factory_b200 import monkeypatch as MP  # noqa: E402
from kvcache_factory_b200 import pyramidkv_utils as U  # noqa: E402
from oracle import kvc_oracle as O  # noqa: E402


@pytest.fixture(autouse=True)
def _restore():
    yield
    MP.restore()


def tiny(cls=LlamaForCausalLM, cfg_cls=LlamaConfig, dtype=torch.float32, head_dim=16, layers=2, **over):
    cfg = cfg_cls(vocab_size=97, hidden_size=4 * head_dim, intermediate_size=64, num_hidden_layers=layers,
                  num_attention_heads=4, num_key_value_heads=2, head_dim=head_dim, max_position_embeddings=4096,
                  attn_implementation="eager", **over)
    torch.manual_seed(0)
    return cls(cfg).to(dtype).eval()


class OracleSnap:
    """SnapKVCluster stand-in on the CPU restatement (test seam: records what the hook passes in)."""
    calls = []

    def __init__(self, window_size=64, max_capacity_prompt=320, kernel_size=5, pooling="avgpool", merge=None, **_):
        self.w, self.cap, self.k, self.pool = window_size, max_capacity_prompt, kernel_size, pooling

    def update_kv(self, key_states, query_states, value_states, attention_mask, num_key_value_groups):
        assert key_states.shape[1] * num_key_value_groups == query_states.shape[1], "hook must pass un-repeated K/V"
        kc, vc, idx, _ = O.snapkv_update_kv(key_states, query_states, value_states, self.w, self.cap, self.k, self.pool)
        OracleSnap.calls.append(dict(L=key_states.shape[-2], out=kc.shape))
        return kc, vc


def test_method_strings_and_restore():
    import transformers.models.llama.modeling_llama as L
    import transformers.models.mistral.modeling_mistral as M
    orig_l, orig_m = L.LlamaAttention.forward, M.MistralAttention.forward
    for m in ("pyramidkv", "snapkv", "h2o", "streamingllm", "l2norm", "think", "adakv", "headkv"):
        kvc.replace_llama(m)
        assert L.LlamaAttention.forward is not orig_l and m in L.LlamaAttention.forward.__name__
        kvc.replace_mistral(m)
        assert M.MistralAttention.forward is not orig_m
    kvc.replace_llama("fullkv", model_name="anything")               # reference: no patch for fullkv
    MP.restore()
    assert L.LlamaAttention.forward is orig_l and M.MistralAttention.forward is orig_m
    kvc.replace_llama("fullkv")
    assert L.LlamaAttention.forward is orig_l
    for m in ("cam", "minference"):
        with pytest.raises(NotImplementedError):
            kvc.replace_llama(m)
    with pytest.raises(ValueError):
        kvc.replace_mistral("nope")


@pytest.mark.parametrize("family", ["llama", "mistral"])
def test_passthrough_equals_plain_transformers(family):
    """Prompt shorter than max_capacity_prompt: update_kv returns its inputs, generation is unchanged."""
    model = tiny() if family == "llama" else tiny(MistralForCausalLM, MistralConfig)
    ids = torch.randint(0, 97, (1, 40), generator=torch.Generator().manual_seed(1))
    ref = model.generate(ids, max_new_tokens=6, do_sample=False, use_cache=True)
    (kvc.replace_llama if family == "llama" else kvc.replace_mistral)("snapkv")
    model.config.max_capacity_prompt, model.config.window_size = 64, 8
    out = model.generate(ids, max_new_tokens=6, do_sample=False, use_cache=True)
    assert torch.equal(out, ref)


def test_compressed_cache_flow_on_cpu(monkeypatch):
    """Prefill stores [B,Hq,cap,d]; every decode step appends one row per query head; RoPE positions are the
    true token positions (generate == a hand-driven loop with explicit position_ids)."""
    monkeypatch.setattr(U, "SnapKVCluster", OracleSnap)
    OracleSnap.calls.clear()
    model = tiny()
    kvc.replace_llama("snapkv")
    model.config.max_capacity_prompt, model.config.window_size, model.config.kernel_size = 24, 8, 5
    L, new = 50, 5
    ids = torch.randint(0, 97, (1, L), generator=torch.Generator().manual_seed(2))
    out = model.generate(ids, max_new_tokens=new, do_sample=False, use_cache=True, return_dict_in_generate=True,
                         output_logits=True)
    assert [c["L"] for c in OracleSnap.calls] == [L, L] and all(c["out"] == (1, 4, 24, 16) for c in OracleSnap.calls)
    cache = out.past_key_values
    # prefill kept 24 rows; the last generated token is never fed back, so new-1 rows were appended
    assert cache.layers[0].keys.shape == (1, 4, 24 + new - 1, 16)
    assert model.model.layers[0].self_attn.kv_seq_len == L + new - 1
    # hand-driven reference: same patched model, explicit positions
    OracleSnap.calls.clear()
    from transformers import DynamicCache
    pkv = DynamicCache(config=model.config)
    with torch.no_grad():
        o = model(input_ids=ids, past_key_values=pkv, use_cache=True)
        logits = [o.logits[:, -1]]
        tok = logits[-1].argmax(-1, keepdim=True)
        for t in range(new - 1):
            o = model(input_ids=tok, position_ids=torch.tensor([[L + t]]), past_key_values=pkv, use_cache=True)
            logits.append(o.logits[:, -1])
            tok = logits[-1].argmax(-1, keepdim=True)
    for a, b in zip(out.logits, logits):
        assert torch.allclose(a, b, atol=1e-5)


class OracleThink:
    """SnapKVCluster stand-in for the ThinK hook: ``update_think`` on the CPU restatement; ``update_kv`` returns the same
    compressed cache with the dropped channels of the non-recent rows ZEROED -- q.k over the kept channels equals q.k_zeroed,
    so the plain compressed-cache hook on this cluster must reproduce the ThinK hook's logits."""

    knobs = (32, 0.4)     # init_snapkv does not pass ThinK's two knobs: the test sets them here for the zeroed-cache run

    def __init__(self, window_size=64, max_capacity_prompt=320, kernel_size=5, pooling="avgpool", merge=None,
                 recent_size=None, ratio=None):
        recent_size = OracleThink.knobs[0] if recent_size is None else recent_size
        ratio = OracleThink.knobs[1] if ratio is None else ratio
        self.kw = dict(window_size=window_size, max_capacity_prompt=max_capacity_prompt, kernel_size=kernel_size,
                       pooling=pooling, recent_size=recent_size, ratio=ratio)
        self.recent_size, self.ratio = recent_size, ratio

    def update_think(self, key_states, query_states, value_states, attention_mask, num_key_value_groups):
        res = O.think_update_kv(key_states, query_states, value_states, **self.kw)
        return res if len(res) == 2 else res[:4]

    def update_kv(self, key_states, query_states, value_states, attention_mask, num_key_value_groups):
        res = O.think_update_kv(key_states, query_states, value_states, **self.kw)
        if len(res) == 2:
            return res
        _, _, mask, vc, kc, _ = res
        kz = kc.clone()
        rows = kc.shape[2] - self.recent_size
        kz[:, :, :rows, :] = kz[:, :, :rows, :] * mask.unsqueeze(2).to(kz.dtype)
        return kz, vc


@pytest.mark.parametrize("ratio", [0.07, 0.4])
def test_think_hook_equals_zeroed_channel_cache_on_cpu(monkeypatch, ratio):
    """ThinK through the hooks: pruned keys + kept-channel mask + recent keys + all values in the cache; decode logits equal
    the ordinary compressed-cache hook on a cache whose dropped channels are zero."""
    monkeypatch.setattr(U, "SnapKVCluster", OracleThink)
    monkeypatch.setattr(OracleThink, "knobs", (6, ratio))
    L, new = 60, 5
    ids = torch.randint(0, 97, (1, L), generator=torch.Generator().manual_seed(4))
    outs = {}
    for method in ("think", "snapkv"):
        MP.restore()
        model = tiny()
        kvc.replace_llama(method)
        c = model.config
        c.max_capacity_prompt, c.window_size, c.kernel_size, c.pooling, c.merge = 24, 8, 5, "avgpool", None
        c.recent_size, c.ratio = 6, ratio
        outs[method] = model.generate(ids, max_new_tokens=new, do_sample=False, use_cache=True,
                                      return_dict_in_generate=True, output_logits=True)
    MP.restore()
    cache = outs["think"].past_key_values
    side = cache.__dict__["_kvc_think"]
    d, k = 16, int(16 * ratio)
    assert side[0][0].shape == (1, 4, 24 - 6, d - k) and side[0][1].shape == (1, 4, d) and side[0][1].dtype == torch.bool
    assert cache.layers[0].keys.shape == (1, 4, 6 + new - 1, d) and cache.layers[0].values.shape == (1, 4, 24 + new - 1, d)
    assert torch.equal(outs["think"].sequences, outs["snapkv"].sequences)
    for a, b in zip(outs["think"].logits, outs["snapkv"].logits):
        assert torch.allclose(a, b, atol=2e-5), float((a - b).abs().max())


def test_think_hook_passthrough_below_capacity():
    model = tiny()
    ids = torch.randint(0, 97, (1, 40), generator=torch.Generator().manual_seed(1))
    ref = model.generate(ids, max_new_tokens=6, do_sample=False, use_cache=True)
    kvc.replace_llama("think")
    model.config.max_capacity_prompt, model.config.window_size = 64, 8
    out = model.generate(ids, max_new_tokens=6, do_sample=False, use_cache=True)
    MP.restore()
    assert torch.equal(out, ref)


@pytest.mark.gpu
def test_snapkv_hook_runs_cuda_path_in_a_tiny_llama():
    dev = torch.device("cuda:0")
    model = tiny(dtype=torch.bfloat16, head_dim=128).to(dev)
    kvc.replace_llama("snapkv")
    w, cap, ks = 32, 128, 7
    model.config.max_capacity_prompt, model.config.window_size, model.config.kernel_size = cap, w, ks
    model.config.pooling = "avgpool"
    captured = []
    orig = U.SnapKVCluster.update_kv

    def spy(self, key_states, query_states, value_states, attention_mask, groups):
        self.keep_debug = True
        out = orig(self, key_states, query_states, value_states, attention_mask, groups)
        captured.append((key_states.detach().cpu(), query_states.detach().cpu(), value_states.detach().cpu(),
                         self.debug["indices"].cpu().long(), out[0].detach().cpu()))
        return out

    U.SnapKVCluster.update_kv = spy
    try:
        L, new = 700, 4
        ids = torch.randint(0, 97, (1, L), generator=torch.Generator().manual_seed(3)).to(dev)
        out = model.generate(ids, max_new_tokens=new, do_sample=False, use_cache=True, return_dict_in_generate=True)
    finally:
        U.SnapKVCluster.update_kv = orig
    torch.cuda.synchronize()
    assert len(captured) == 2
    for K, Q, V, idx, Kc in captured:
        assert K.shape == (1, 2, L, 128) and Q.shape == (1, 4, L, 128)      # un-repeated K/V reach the kernels
        _, _, idx_ref, pooled_ref = O.snapkv_update_kv(K, Q, V, w, cap, ks, "avgpool")
        ok, msg = O.check_index_sets(pooled_ref, idx, idx_ref.shape[-1], rtol=2.0 ** -6)
        assert ok, msg
        Kr = O.repeat_kv(K, 2)
        want = torch.cat([Kr[:, :, :-w].gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, 128)), Kr[:, :, -w:]], 2)
        assert torch.equal(Kc, want)
    cache = out.past_key_values
    assert cache.layers[0].keys.shape == (1, 4, cap + new - 1, 128)
    assert out.sequences.shape == (1, L + new)


@pytest.mark.gpu
def test_think_hook_runs_cuda_path_in_a_tiny_llama(monkeypatch):
    """replace_llama("think") on the GPU: kvc_think_prune feeds the cache; decode logits agree with the plain
    compressed-cache hook run on the same CUDA cluster with the dropped channels zeroed."""
    dev = torch.device("cuda:0")
    w, cap, ks, recent, ratio = 32, 128, 7, 16, 0.4

    class ZeroedSnap(U.SnapKVCluster):
        def update_kv(self, key_states, query_states, value_states, attention_mask, groups, **kw):
            kc, vc = super().update_kv(key_states, query_states, value_states, attention_mask, groups)
            if kc.shape[2] != cap:
                return kc, vc
            _, _, mask = kvc._native.think_prune(kc, query_states, recent, ratio)
            kz = kc.clone()
            kz[:, :, :cap - recent] = kz[:, :, :cap - recent] * mask.unsqueeze(2).to(kz.dtype)
            return kz, vc

    L, new = 700, 4
    ids = torch.randint(0, 97, (1, L), generator=torch.Generator().manual_seed(3)).to(dev)
    outs = {}
    for method in ("think", "snapkv"):
        MP.restore()
        if method == "snapkv":
            monkeypatch.setattr(U, "SnapKVCluster", ZeroedSnap)
        model = tiny(dtype=torch.bfloat16, head_dim=128).to(dev)
        kvc.replace_llama(method)
        c = model.config
        c.max_capacity_prompt, c.window_size, c.kernel_size, c.pooling, c.merge = cap, w, ks, "avgpool", None
        c.recent_size, c.ratio = recent, ratio
        outs[method] = model.generate(ids, max_new_tokens=new, do_sample=False, use_cache=True,
                                      return_dict_in_generate=True, output_logits=True)
    MP.restore()
    torch.cuda.synchronize()
    cache = outs["think"].past_key_values
    pruned, mask = cache.__dict__["_kvc_think"][0]
    k = int(128 * ratio)
    assert pruned.shape == (1, 4, cap - recent, 128 - k) and pruned.is_cuda and mask.dtype == torch.bool
    assert bool((mask.sum(-1) == 128 - k).all())
    assert cache.layers[0].keys.shape == (1, 4, recent + new - 1, 128)
    assert cache.layers[0].values.shape == (1, 4, cap + new - 1, 128)
    for a, b in zip(outs["think"].logits, outs["snapkv"].logits):
        assert bool(torch.isfinite(a).all())
        assert float((a.float() - b.float()).abs().max()) <= 0.05 * float(b.float().abs().max()) + 1e-3


@pytest.mark.gpu
def test_adakv_hook_flattened_cache_and_varlen_decode():
    pytest.importorskip("flash_attn")
    dev = torch.device("cuda:0")
    model = tiny(dtype=torch.bfloat16, head_dim=128).to(dev)
    kvc.replace_llama("adakv")
    w, cap = 32, 128
    c = model.config
    c.max_capacity_prompt, c.window_size, c.kernel_size, c.pooling, c.floor, c.normalize = cap, w, 7, "maxpool", 0.2, True
    L, new = 700, 4
    ids = torch.randint(0, 97, (1, L), generator=torch.Generator().manual_seed(4)).to(dev)
    out = model.generate(ids, max_new_tokens=new, do_sample=False, use_cache=True, return_dict_in_generate=True,
                         output_logits=True)
    torch.cuda.synchronize()
    assert out.sequences.shape == (1, L + new) and all(torch.isfinite(l.float()).all() for l in out.logits)
    flat = out.past_key_values._kvc_flat
    for layer in model.model.layers:
        a = layer.self_attn
        cl = a.kv_cluster
        assert a.kv_seq_len == L + new - 1
        # per-head lengths: budget + window at prefill, +1 per decode step; the flattened tensors hold exactly that
        lens = cl.head_lens.cpu()
        assert int(lens.sum()) == cl.klen_sum == flat.key_cache[a.layer_idx].shape[0]
        # AdaKV reallocates the H*base budget across heads; round() per head may move the total by < 1 row per head
        assert abs(int(lens.sum()) - (4 * (cap - w) + 4 * w + 4 * (new - 1))) <= 4
        assert int(lens.min()) >= int((cap - w) * 0.2) + w + new - 1
        assert torch.equal(cl.cu_klen.cpu()[1:] - cl.cu_klen.cpu()[:-1], lens)
        assert cl.max_seqlen_k >= int(lens.max())


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["snapkv", "pyramidkv"])
def test_async_knobs_through_the_hooks_equal_the_synchronous_path(method, monkeypatch):
    """config.kvc_async_tail / kvc_async_scores are read by init_* on every forward (the clusters are rebuilt each
    time), the hook attaches the still-being-gathered rows to the HF cache without reading them and joins the side
    streams before the first decode read: same tokens, same cache, bit for bit."""
    dev = torch.device("cuda:0")
    # the asynchronous scoring path runs the half-footprint kernel; pin it for the synchronous runs too so that the
    # three runs are bit-comparable (the kernels differ in the grouping of the fp32 partial sums)
    monkeypatch.setattr(U._ScoredCluster, "score_impl", kvc._native.IMPL_TCGEN05_HALF)
    model = tiny(dtype=torch.bfloat16, head_dim=128, layers=4).to(dev)
    kvc.replace_llama(method)
    model.config.max_capacity_prompt, model.config.window_size, model.config.kernel_size = 256, 32, 7
    model.config.pooling = "avgpool"
    ids = torch.randint(0, 97, (1, 3000), generator=torch.Generator().manual_seed(5)).to(dev)
    runs = []
    for flags in ((False, False), (True, False), (True, True)):
        model.config.kvc_async_tail, model.config.kvc_async_scores = flags
        out = model.generate(ids, max_new_tokens=6, do_sample=False, use_cache=True, return_dict_in_generate=True)
        torch.cuda.synchronize()
        assert isinstance(model.model.layers[0].self_attn.kv_cluster.async_tail, bool)
        assert model.model.layers[0].self_attn.kv_cluster.async_tail == flags[0]
        runs.append((out.sequences.cpu(), [l.keys.cpu() for l in out.past_key_values.layers],
                     [l.values.cpu() for l in out.past_key_values.layers]))
    for seq, ks, vs in runs[1:]:
        assert torch.equal(seq, runs[0][0])
        for a, b in zip(ks, runs[0][1]):
            assert torch.equal(a, b)
        for a, b in zip(vs, runs[0][2]):
            assert torch.equal(a, b)


def test_hooks_refuse_a_cache_without_layers():
    class OldCache:                      # transformers < 4.54 style: no per-layer view
        key_cache, value_cache = [], []
    with pytest.raises(TypeError, match="layered cache"):
        MP._layer_keys(OldCache(), 0)


@pytest.mark.gpu
def test_adakv_hook_decode_slack_equals_rebuild_path():
    """config.kvc_decode_slack: the decode steps of AdaKV write their rows in place (re-layout every `slack` steps)
    and attend over (segment starts, used lengths); same tokens and logits as the reference-style rebuild path."""
    pytest.importorskip("flash_attn")
    dev = torch.device("cuda:0")
    model = tiny(dtype=torch.bfloat16, head_dim=128).to(dev)
    kvc.replace_llama("adakv")
    c = model.config
    c.max_capacity_prompt, c.window_size, c.kernel_size, c.pooling, c.floor, c.normalize = 128, 32, 7, "maxpool", 0.2, True
    ids = torch.randint(0, 97, (1, 700), generator=torch.Generator().manual_seed(4)).to(dev)
    outs = []
    for slack in (0, 3):
        c.kvc_decode_slack = slack
        for layer in model.model.layers:            # AdaKV clusters are built once per module: start afresh
            if hasattr(layer.self_attn, "kv_cluster"):
                del layer.self_attn.kv_cluster
        out = model.generate(ids, max_new_tokens=11, do_sample=False, use_cache=True, return_dict_in_generate=True,
                             output_logits=True)
        torch.cuda.synchronize()
        outs.append(out)
        assert out.past_key_values._kvc_flat.slack == slack
    assert torch.equal(outs[0].sequences, outs[1].sequences)
    for a, b in zip(outs[0].logits, outs[1].logits):
        assert torch.equal(a, b)
    flat = outs[1].past_key_values._kvc_flat
    lens = model.model.layers[0].self_attn.kv_cluster.head_lens
    assert int(flat.cu_cap[0][-1]) == flat.key_cache[0].shape[0] >= int(lens.sum())


@pytest.mark.gpu
def test_scores_from_lse_through_the_hooks():
    """config.kvc_scores_from_lse: the prefill attention runs through flash attention, whose log-sum-exp replaces the
    first softmax pass of the scoring kernel; the cache has the same shape and (nearly) the same rows."""
    pytest.importorskip("flash_attn")
    dev = torch.device("cuda:0")
    model = tiny(dtype=torch.bfloat16, head_dim=128, layers=2).to(dev)
    kvc.replace_llama("snapkv")
    model.config.max_capacity_prompt, model.config.window_size, model.config.kernel_size = 256, 32, 7
    model.config.pooling = "avgpool"
    ids = torch.randint(0, 97, (1, 2000), generator=torch.Generator().manual_seed(6)).to(dev)
    caches = []
    for flag in (False, True):
        model.config.kvc_scores_from_lse = flag
        out = model.generate(ids, max_new_tokens=3, do_sample=False, use_cache=True, return_dict_in_generate=True)
        torch.cuda.synchronize()
        caches.append(out.past_key_values.layers[0].keys[:, :, :256].cpu())
        assert out.sequences.shape == (1, 2003)
    a, b = caches
    assert a.shape == b.shape == (1, 4, 256, 128)
    same_rows = np.mean([len({tuple(r.tolist()[:4]) for r in a[0, h]} & {tuple(r.tolist()[:4]) for r in b[0, h]}) / 256
                         for h in range(4)])
    assert same_rows > 0.9, same_rows
