"""CPU-side tests: the C-ABI library builds, loads and exports every symbol include/kvc_b200.h
declares; the host-side mirror of the reference interface (budgets, defaults, errors) behaves like
the reference.  No compute calls: there is no GPU here and no CPU fallback to call."""
import ctypes
import os
import re
import types

import pytest
import torch

import kvcache_factory_b200 as kvc
from kvcache_factory_b200 import _native
from oracle import kvc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as G
        G.build()
    return _native.lib()


def test_header_symbols_are_all_exported(built_lib):
    header = open(os.path.join(ROOT, "include", "kvc_b200.h")).read()
    declared = set(re.findall(r"\b(kvc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/kvc_b200.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert built_lib.kvc_version() == _native.ABI_VERSION


def test_library_is_sm100a_only(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_workspace_queries_need_no_gpu(built_lib):
    assert built_lib.kvc_pool_topk_workspace_bytes(32, 32736, 3932) >= 32 * 32736 * 2
    assert built_lib.kvc_adakv_allocate_workspace_bytes(32, 65504) > 32 * 65504 * 2
    assert built_lib.kvc_pool_topk_workspace_bytes(0, 10, 1) == 0


@pytest.mark.parametrize("L,w,cap,layers,beta", [(32768, 32, 2048, 32, 20), (131072, 32, 2048, 80, 20),
                                                 (5000, 8, 2048, 32, 20), (4096, 32, 128, 32, 20),
                                                 (700, 16, 320, 12, 10), (2049, 64, 1024, 40, 20)])
def test_pyramid_budget_equals_oracle(L, w, cap, layers, beta):
    for layer in range(layers):
        assert kvc.pyramid_budget(L, w, cap, layers, layer, beta) == O.pyramidkv_budget(L, w, cap, layers, layer, beta)
    cl = kvc.PyramidKVCluster(num_hidden_layers=layers, window_size=w, max_capacity_prompt=cap, beta=beta,
                              layer_idx=layers - 1)
    assert cl.layer_budget(L) == O.pyramidkv_budget(L, w, cap, layers, layers - 1, beta)


def test_constructor_contract():
    # same defaults as the reference (pyramidkv_utils.py:174,286,517,579,626)
    s = kvc.SnapKVCluster()
    assert (s.window_size, s.max_capacity_prompt, s.kernel_size, s.pooling, s.merge) == (64, 320, 5, "avgpool", None)
    p = kvc.PyramidKVCluster()
    assert (p.num_hidden_layers, p.beta, p.layer_idx, p.steps) == (32, 20, None, -1)
    a = kvc.AdaKVCluster(max_capacity_prompt=2048, floor=0.2, normalize=True)
    assert (a.window_size, a.kernel_size, a.pooling) == (32, 7, "maxpool")
    assert (a.base_capacity, a.floor_capacity, a.adaptive_capacity) == (2016, 403, 1613)
    for cls in (kvc.SnapKVCluster, kvc.PyramidKVCluster, kvc.H2OKVCluster, kvc.StreamingLLMKVCluster):
        with pytest.raises(AssertionError):
            cls(window_size=64, max_capacity_prompt=64)
        c = cls(window_size=8, max_capacity_prompt=64)
        c.reset(window_size=16, max_capacity_prompt=128, kernel_size=7, pooling="maxpool")
        assert (c.window_size, c.max_capacity_prompt, c.kernel_size, c.pooling) == (16, 128, 7, "maxpool")
        with pytest.raises(AssertionError):
            c.reset(window_size=64, max_capacity_prompt=64)


def test_no_cpu_fallback():
    Q = torch.zeros(1, 4, 256, 128, dtype=torch.bfloat16)
    K = torch.zeros(1, 2, 256, 128, dtype=torch.bfloat16)
    for cl in (kvc.SnapKVCluster(16, 64), kvc.PyramidKVCluster(window_size=16, max_capacity_prompt=64, layer_idx=0),
               kvc.H2OKVCluster(16, 64), kvc.StreamingLLMKVCluster(16, 64)):
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            cl.update_kv(K, Q, K, None, 2)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        kvc.AdaKVCluster(window_size=16, max_capacity_prompt=64, floor=0.2, normalize=True).update_kv(K, Q, K)
    # pass-through needs no device work and returns the same objects (pyramidkv_utils.py:314-315)
    k2, v2 = kvc.SnapKVCluster(16, 512).update_kv(K, Q, K, None, 2)
    assert k2 is K and v2 is K
    with pytest.raises(AssertionError):
        kvc.SnapKVCluster(16, 64).update_kv(K[:, :, :100], Q, K, None, 2)
    with pytest.raises(ValueError):
        kvc.SnapKVCluster(16, 64).update_kv(K, Q[:, :3], K, None, 2)      # 2 KV heads cannot serve 3 query heads
    with pytest.raises(NotImplementedError):
        kvc.SnapKVCluster(16, 64, merge="pivot").update_kv(K, Q, K, None, 2)


def _attn_stub(layer_idx=3, **cfg):
    return types.SimpleNamespace(layer_idx=layer_idx, config=types.SimpleNamespace(num_hidden_layers=32, **cfg))


def test_init_helpers_fill_reference_defaults():
    m = _attn_stub()
    kvc.init_snapkv(m)
    assert isinstance(m.kv_cluster, kvc.SnapKVCluster)
    assert (m.config.window_size, m.config.max_capacity_prompt, m.config.kernel_size, m.config.pooling) == \
        (32, 4096, 5, "avgpool")                                   # pyramidkv_utils.py:904-915
    first = m.kv_cluster
    kvc.init_snapkv(m)
    assert m.kv_cluster is not first                               # rebuilt on every call (:918)

    m = _attn_stub(max_capacity_prompt=512, window_size=8)
    kvc.init_pyramidkv(m, num_hidden_layers=32)
    assert (m.kv_cluster.max_capacity_prompt, m.kv_cluster.window_size, m.kv_cluster.layer_idx) == (512, 8, 3)
    assert m.config.kernel_size == 5 and m.config.pooling == "avgpool"

    for init, cls in ((kvc.init_H2O, kvc.H2OKVCluster), (kvc.init_StreamingLLM, kvc.StreamingLLMKVCluster)):
        m = _attn_stub()
        init(m)
        assert isinstance(m.kv_cluster, cls) and m.config.max_capacity_prompt == 2048

    m = _attn_stub()
    with pytest.raises(AttributeError):                            # reads config.floor, defaults floor_ratio (:1057)
        kvc.init_adakv(m)
    m = _attn_stub(floor=0.2)
    kvc.init_adakv(m)
    assert (m.config.pooling, m.config.normalize, m.config.floor_ratio) == ("maxpool", True, 0.2)
    first = m.kv_cluster
    kvc.init_adakv(m)
    assert m.kv_cluster is first                                   # built once (:1049)


def test_flatten_cache_container_protocol():
    c = kvc.DynamicCacheSplitHeadFlatten()
    assert len(c) == 0 and c.get_seq_length(0) == 0 and c.get_max_length() is None
    k = torch.zeros(10, 128)
    c.update(k, k, 0)                                              # prefill append needs no kernel
    assert len(c) == 1 and c.get_seq_length(0) == 1 and c[0][0][0].shape == (128,)
    with pytest.raises(KeyError):
        c[1]
    legacy = c.to_legacy_cache()
    c2 = kvc.DynamicCacheSplitHeadFlatten.from_legacy_cache(legacy)
    assert len(c2) == 1 and c2.key_cache[0] is k
