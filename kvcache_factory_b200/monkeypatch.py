"""``replace_llama`` / ``replace_mistral``: the reference's model hooks (``pyramidkv/monkeypatch.py:19-145``)
for the ``transformers`` release installed today.

The reference patches ``LlamaAttention`` / ``LlamaFlashAttention2`` / ``LlamaSdpaAttention.forward`` with one
hand-copied forward per (method, attention backend) pair (``pyramidkv/llama_model.py``, ``mistral_model.py``)
written against transformers 4.44; those classes no longer exist (one ``LlamaAttention`` dispatches through
``ALL_ATTENTION_FUNCTIONS``), so the reference's layer cannot be imported any more.  This module keeps the
hook NAMES, SIGNATURES and METHOD STRINGS and installs ONE forward per model family that

* prefill (the layer's cache is empty): attention over the full prompt through HF's own attention
  interface, then ``self.kv_cluster.update_kv(key_states, query_states, value_states, attention_mask,
  self.num_key_value_groups)`` -- handed the UN-repeated K/V (the kernels fold the GQA groups, the
  reference materialises ``repeat_kv`` first, ``llama_model.py:379-380``) -- and stores the compressed
  ``[B, Hq, cap, d]`` rows in the HF cache (``llama_model.py:399-404``);
* decode: the new K/V row is expanded to the cache's head count, appended, and the query attends to the
  compressed cache without a mask, as in the reference (``llama_model.py:405-445``); ``self.kv_seq_len``
  tracks the true sequence length exactly like the reference does.

Positions: ``generate`` derives ``position_ids`` from the attention mask, so RoPE keeps the true token
positions even though the cache is shorter than the prompt (the reference overrides
``prepare_inputs_for_generation`` for the same purpose).  Callers that drive ``model.forward`` by hand must
pass ``position_ids`` themselves once a cache is compressed.

AdaKV / HeadKV keep per-head lengths in the flattened ``[sum_h len_h, d]`` layout
(``DynamicCacheSplitHeadFlatten``); their decode step needs ``flash_attn_varlen_func``
(``llama_model.py:2362-2389``), imported lazily.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional

import torch

from . import pyramidkv_utils as U

DENSE_METHODS = ("pyramidkv", "snapkv", "h2o", "streamingllm", "l2norm")
FLAT_METHODS = ("adakv", "headkv")
# methods of the reference that lie outside the hot path rebuilt here ("out of scope")
THINK_METHODS = ("think",)
UNSUPPORTED = ("cam", "minference")

_ORIGINAL: Dict[type, Callable] = {}


def _init_cluster(module, method: str) -> None:
    """What every reference forward does first (``init_pyramidkv(self, num_hidden_layers=...)`` etc.)."""
    if method == "pyramidkv":
        U.init_pyramidkv(module, num_hidden_layers=module.config.num_hidden_layers)
    elif method == "snapkv":
        U.init_snapkv(module)
    elif method == "h2o":
        U.init_H2O(module)
    elif method == "streamingllm":
        U.init_StreamingLLM(module)
    elif method == "l2norm":
        U.init_l2norm(module)
    elif method == "think":
        U.init_think(module)
    elif method == "adakv":
        U.init_adakv(module)
    elif method == "headkv":
        U.init_headkv(module)


class _GroupsView:
    """The attention module as seen by HF's attention functions, with a different GQA group count
    (the compressed cache holds one K/V head per QUERY head)."""

    def __init__(self, module, groups: int):
        object.__setattr__(self, "_module", module)
        object.__setattr__(self, "num_key_value_groups", groups)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_module"), name)


def _layer_keys(past_key_values, layer_idx: int) -> Optional[torch.Tensor]:
    layers = getattr(past_key_values, "layers", None)
    if layers is None:
        # Without the per-layer view there is no way to tell prefill from decode: every decode step would be
        # compressed as a one-token "prefill" and attention would silently see only the current token.
        raise TypeError(f"kvcache_factory_b200 hooks need a transformers >= 4.54 layered cache (DynamicCache with "
                        f".layers); got {type(past_key_values).__name__}")
    if layer_idx >= len(layers):
        return None
    keys = getattr(layers[layer_idx], "keys", None)
    if keys is None or not getattr(layers[layer_idx], "is_initialized", True) or keys.numel() == 0:
        return None
    return keys


def _store_prefill(past_key_values, kc: torch.Tensor, vc: torch.Tensor, layer_idx: int, pending: bool) -> None:
    """Hand the compressed rows to the HF cache (``llama_model.py:399-404``).  ``DynamicLayer.update`` concatenates
    onto an empty tensor, i.e. READS the rows on the caller's stream; with ``async_tail`` they are still being
    gathered on a side stream, so the layer is created with empty slices (no data is touched) and the tensors are
    attached as they are -- which also saves a copy of the compressed cache in the synchronous case.  The first decode
    step joins the side streams (``_join_compression``)."""
    if kc.dim() == 4 and hasattr(past_key_values, "layers"):
        past_key_values.update(kc[:, :, :0], vc[:, :, :0], layer_idx)
        layer = past_key_values.layers[layer_idx]
        # plain DynamicLayer only: sliding-window / quantised layers keep their own bookkeeping in update()
        if type(layer).__name__ == "DynamicLayer" and layer.keys.shape[-2] == 0:
            layer.keys, layer.values = kc, vc
            if pending:
                past_key_values.__dict__["_kvc_pending"] = True
            return
    if pending:
        U.wait_for_compression(kc.device)
    past_key_values.update(kc, vc, layer_idx)


def _join_compression(past_key_values, device) -> None:
    """Before the compressed cache is read for the first time: wait for the asynchronous select/gather tails."""
    if past_key_values.__dict__.get("_kvc_pending"):
        U.wait_for_compression(device)
        past_key_values.__dict__["_kvc_pending"] = False


def _make_forward(method: str, family):
    apply_rotary_pos_emb = family.apply_rotary_pos_emb
    eager_attention_forward = family.eager_attention_forward
    functions = family.ALL_ATTENTION_FUNCTIONS
    repeat_kv = family.repeat_kv

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        q_len = query_states.shape[-2]

        interface = functions.get_interface(self.config._attn_implementation, eager_attention_forward)
        extra = dict(kwargs)
        if hasattr(self.config, "sliding_window") and "sliding_window" not in extra and family.__name__.endswith("mistral"):
            extra["sliding_window"] = getattr(self.config, "sliding_window", None)
        common = dict(dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling)

        module_for_attention, mask = self, attention_mask
        if past_key_values is None:                                   # no cache: nothing to compress
            k_att, v_att = key_states, value_states
        elif _layer_keys(past_key_values, self.layer_idx) is None:    # ---- prefill ----
            _init_cluster(self, method)
            self.kv_seq_len = q_len                                   # llama_model.py:362-376
            k_att, v_att = key_states, value_states                   # attention sees the whole prompt
            lse_kw = {}
            if method in ("snapkv", "pyramidkv") and getattr(self.config, "kvc_scores_from_lse", False) \
                    and query_states.is_cuda and query_states.dtype in (torch.bfloat16, torch.float16) \
                    and q_len >= getattr(self.kv_cluster, "max_capacity_prompt", 1 << 62):
                # SURVEY.md section 8 f4 (opt-in, config.kvc_scores_from_lse): the prefill runs through flash attention
                # directly, which hands back the log-sum-exp of every query row; the observation window's rows replace
                # the first softmax pass of the scoring kernel
                from flash_attn import flash_attn_func
                fa_out, fa_lse, _ = flash_attn_func(query_states.transpose(1, 2), key_states.transpose(1, 2),
                                                    value_states.transpose(1, 2), dropout_p=0.0, softmax_scale=self.scaling,
                                                    causal=True, return_attn_probs=True)
                lse_kw = dict(window_lse=fa_lse[:, :, -self.kv_cluster.window_size:].float().contiguous())
                kc, vc = self.kv_cluster.update_kv(key_states, query_states, value_states, attention_mask,
                                                   self.num_key_value_groups, **lse_kw)
                _store_prefill(past_key_values, kc, vc, self.layer_idx, getattr(self.kv_cluster, "async_tail", False))
                return self.o_proj(fa_out.reshape(*input_shape, -1).contiguous()), None
            kc, vc = self.kv_cluster.update_kv(key_states, query_states, value_states, attention_mask,
                                               self.num_key_value_groups)
            _store_prefill(past_key_values, kc, vc, self.layer_idx, getattr(self.kv_cluster, "async_tail", False))
        else:                                                         # ---- decode ----
            _join_compression(past_key_values, hidden_states.device)
            self.kv_seq_len = getattr(self, "kv_seq_len", 0) + q_len
            cached = _layer_keys(past_key_values, self.layer_idx)
            if cached.shape[1] != key_states.shape[1]:                # compressed cache: one K/V head per query head
                key_states = repeat_kv(key_states, self.num_key_value_groups)
                value_states = repeat_kv(value_states, self.num_key_value_groups)
            k_att, v_att = past_key_values.update(key_states, value_states, self.layer_idx)
            groups = query_states.shape[1] // k_att.shape[1]
            if groups != self.num_key_value_groups:
                module_for_attention = _GroupsView(self, groups)
            if k_att.shape[-2] != self.kv_seq_len:                    # rows were evicted: the prompt-shaped mask no longer applies
                mask = None
        attn_output, attn_weights = interface(module_for_attention, query_states, k_att, v_att, mask, **common, **extra)
        attn_output = attn_output.reshape(*input_shape, -1).contiguous()
        return self.o_proj(attn_output), attn_weights

    forward.__name__ = f"kvc_{method}_attention_forward"
    return forward


def _make_flat_forward(method: str, family):
    """AdaKV / HeadKV (``llama_flash_attn2_forward_AdaKV`` / ``_HeadKV``, ``llama_model.py:2229-2393``)."""
    apply_rotary_pos_emb = family.apply_rotary_pos_emb
    eager_attention_forward = family.eager_attention_forward
    functions = family.ALL_ATTENTION_FUNCTIONS

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        bsz, q_len = input_shape
        hidden_shape = (*input_shape, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        num_heads = query_states.shape[1]
        if past_key_values is None:
            raise ValueError(f"{method} needs use_cache=True: its decode path reads the flattened per-head cache")
        # The reference swaps in DynamicCacheSplitHeadFlatten through a patched LlamaModel.forward
        # (adaptive_LlamaModel_forward, llama_model.py:2420-2430); here the flattened store rides on whatever
        # cache object HF created for this generate() call.
        store = past_key_values
        if not isinstance(store, U.DynamicCacheSplitHeadFlatten):
            # config.kvc_decode_slack (or KVC_DECODE_SLACK): rows of slack per head segment -> O(H*d) decode appends
            slack = int(getattr(self.config, "kvc_decode_slack", os.environ.get("KVC_DECODE_SLACK", "0")))
            store = past_key_values.__dict__.setdefault("_kvc_flat", U.DynamicCacheSplitHeadFlatten(slack=slack))
        past_key_values = store
        if len(past_key_values) <= self.layer_idx:                    # ---- prefill ----
            _init_cluster(self, method)
            self.kv_seq_len = q_len
            interface = functions.get_interface(self.config._attn_implementation, eager_attention_forward)
            attn_output, attn_weights = interface(self, query_states, key_states, value_states, attention_mask,
                                                  dropout=0.0, scaling=self.scaling, **kwargs)
            flat_k, flat_v = self.kv_cluster.update_kv(key_states, query_states, value_states)
            past_key_values.update(flat_k, flat_v, self.layer_idx,
                                   {"head_lens": self.kv_cluster.head_lens, "cu_klen": self.kv_cluster.cu_klen})
            attn_output = attn_output.reshape(*input_shape, -1).contiguous()
        else:                                                         # ---- decode (llama_model.py:2362-2389) ----
            from flash_attn import flash_attn_varlen_func
            assert bsz == 1 and q_len == 1
            self.kv_seq_len += q_len
            c = self.kv_cluster
            groups = num_heads // key_states.shape[1]
            k_new = U.repeat_kv(key_states, groups)                   # the flattened cache has one segment per query head
            v_new = U.repeat_kv(value_states, groups)
            flat_k, flat_v = past_key_values.update(k_new, v_new, self.layer_idx,
                                                    {"head_lens": c.head_lens, "cu_klen": c.cu_klen})
            c.klen_sum += num_heads
            c.max_seqlen_k += 1
            c.head_lens += 1
            q = query_states.reshape(-1, 1, self.head_dim)
            cu_cap = past_key_values.cu_cap[self.layer_idx] if past_key_values.slack > 0 else None
            if cu_cap is not None:
                # slack layout: segment STARTS + USED lengths instead of packed cu_seqlens (the rows were written in place)
                from flash_attn.flash_attn_interface import _flash_attn_varlen_forward
                attn_output = _flash_attn_varlen_forward(
                    q, flat_k.view(-1, 1, self.head_dim), flat_v.view(-1, 1, self.head_dim), c.cu_qlen, cu_cap, 1,
                    c.max_seqlen_k, 0.0, self.head_dim ** -0.5, True, seqused_k=c.head_lens)[0]
            else:
                c.cu_klen += c.cu_offset
                attn_output = flash_attn_varlen_func(q, flat_k.view(-1, 1, self.head_dim), flat_v.view(-1, 1, self.head_dim),
                                                     c.cu_qlen, c.cu_klen, 1, c.max_seqlen_k, causal=True)
            attn_output = attn_output.reshape(bsz, num_heads, q_len, self.head_dim).transpose(1, 2)
            attn_output = attn_output.reshape(bsz, q_len, -1).contiguous()
            attn_weights = None
        past_key_values._seen_tokens = self.kv_seq_len
        return self.o_proj(attn_output), attn_weights

    forward.__name__ = f"kvc_{method}_attention_forward"
    return forward


def _make_think_forward(method: str, family):
    """ThinK (``llama_attn_forward_think``, ``llama_model_think.py:120-216``; cache ``cache_utils_think.py:390-424``):
    the prefill stores the channel-pruned keys, the kept-channel mask, the unpruned recent keys and the compressed values;
    a decode step scores the pruned keys with the query's kept channels and the recent + new keys with the whole query."""
    apply_rotary_pos_emb = family.apply_rotary_pos_emb
    eager_attention_forward = family.eager_attention_forward
    functions = family.ALL_ATTENTION_FUNCTIONS
    repeat_kv = family.repeat_kv

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        bsz, num_heads, q_len, d = query_states.shape
        if past_key_values is None:
            raise ValueError("think needs use_cache=True: its decode path reads the pruned key cache")
        side = past_key_values.__dict__.setdefault("_kvc_think", {})
        if _layer_keys(past_key_values, self.layer_idx) is None:      # ---- prefill ----
            _init_cluster(self, method)
            if self.kv_cluster.recent_size <= 0:
                raise ValueError("think: recent_size must be positive (the unpruned recent keys mark the layer as prefilled)")
            self.kv_seq_len = q_len
            interface = functions.get_interface(self.config._attn_implementation, eager_attention_forward)
            attn_output, attn_weights = interface(self, query_states, key_states, value_states, attention_mask,
                                                  dropout=0.0 if not self.training else self.attention_dropout,
                                                  scaling=self.scaling, **kwargs)
            res = self.kv_cluster.update_think(key_states, query_states, value_states, attention_mask,
                                               self.num_key_value_groups)
            if len(res) == 2:                                         # below the capacity: nothing pruned (:357-358)
                side[self.layer_idx] = None
                past_key_values.update(repeat_kv(res[0], num_heads // res[0].shape[1]),
                                       repeat_kv(res[1], num_heads // res[1].shape[1]), self.layer_idx)
            else:
                kv_pruned, kv_recent, mask, vc = res
                side[self.layer_idx] = (kv_pruned.view(bsz, num_heads, -1, kv_pruned.shape[-1]), mask)
                _store_prefill(past_key_values, kv_recent, vc, self.layer_idx, False)   # keys: recent rows, values: all
            return self.o_proj(attn_output.reshape(*input_shape, -1).contiguous()), attn_weights
        # ---- decode (llama_model_think.py:176-199) ----
        if q_len != 1:
            raise NotImplementedError("think: the pruned-key attention of the reference is defined for one query token")
        self.kv_seq_len = getattr(self, "kv_seq_len", 0) + q_len
        key_states = repeat_kv(key_states, self.num_key_value_groups)
        value_states = repeat_kv(value_states, self.num_key_value_groups)
        layer = past_key_values.layers[self.layer_idx]
        if type(layer).__name__ == "DynamicLayer":                    # keys and values have different lengths here
            layer.keys = torch.cat([layer.keys, key_states], dim=-2)
            layer.values = torch.cat([layer.values, value_states], dim=-2)
            k_recent, v_all = layer.keys, layer.values
        else:
            k_recent, v_all = past_key_values.update(key_states, value_states, self.layer_idx)
        scale = d ** -0.5
        w2 = torch.matmul(query_states, k_recent.transpose(2, 3)) * scale
        state = side.get(self.layer_idx)
        if state is not None:
            key_pruned, mask = state
            mq = query_states[mask.unsqueeze(2).expand(-1, -1, q_len, -1)].view(bsz, num_heads, q_len, -1)
            w1 = torch.matmul(mq, key_pruned.transpose(2, 3)) * scale
            w2 = torch.cat([w1, w2], dim=-1)
        probs = torch.softmax(w2, dim=-1, dtype=torch.float32).to(query_states.dtype)
        attn_output = torch.matmul(probs, v_all).transpose(1, 2)
        return self.o_proj(attn_output.reshape(*input_shape, -1).contiguous()), None

    forward.__name__ = "kvc_think_attention_forward"
    return forward


def _replace(method: str, family, attention_cls) -> None:
    if method == "fullkv":
        return
    if method in UNSUPPORTED:
        raise NotImplementedError(f"method '{method}' is outside the compression hot path rebuilt here "
                                  f"(supported: {', '.join(DENSE_METHODS + FLAT_METHODS + THINK_METHODS)}, fullkv)")
    if method not in DENSE_METHODS + FLAT_METHODS + THINK_METHODS:
        raise ValueError(f"unknown method '{method}'")
    _ORIGINAL.setdefault(attention_cls, attention_cls.forward)
    maker = _make_forward if method in DENSE_METHODS else (_make_think_forward if method in THINK_METHODS else _make_flat_forward)
    attention_cls.forward = maker(method, family)


def replace_llama(method, model_name=None):
    """Reference ``replace_llama(method, model_name=None)`` (``monkeypatch.py:19-83``)."""
    import transformers.models.llama.modeling_llama as family
    _replace(method, family, family.LlamaAttention)


def replace_mistral(method):
    """Reference ``replace_mistral(method)`` (``monkeypatch.py:88-145``)."""
    import transformers.models.mistral.modeling_mistral as family
    _replace(method, family, family.MistralAttention)


def restore() -> None:
    """Undo every ``replace_*`` (not in the reference; used by the tests)."""
    for cls, fwd in _ORIGINAL.items():
        cls.forward = fwd
    _ORIGINAL.clear()
