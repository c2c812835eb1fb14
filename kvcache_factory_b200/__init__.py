"""kvcache_factory_b200 — B200-native (sm_100a) prefill KV-cache compression.

Drop-in for the compression policies of assassin808/KVCache-Factory's
``pyramidkv/pyramidkv_utils.py`` (SnapKV, PyramidKV, H2O, StreamingLLM, L2Norm, AdaKV/HeadKV and the flattened
AdaKV cache append).  The Python classes keep the reference's names and signatures; the work is
done by hand-written CUDA kernels in ``libkvc_b200.so`` behind a C ABI (``include/kvc_b200.h``).
Importing the package does not need a GPU; calling any ``update_kv`` does.
"""
from .monkeypatch import replace_llama, replace_mistral  # noqa: F401
from .pyramidkv_utils import (  # noqa: F401
    AdaKVCluster,
    DynamicCacheSplitHeadFlatten,
    H2OKVCluster,
    HeadKVCluster,
    L2NormCluster,
    PyramidKVCluster,
    SnapKVCluster,
    StreamingLLMKVCluster,
    init_adakv,
    init_headkv,
    init_l2norm,
    init_think,
    init_H2O,
    init_pyramidkv,
    init_snapkv,
    init_StreamingLLM,
    pyramid_budget,
    repeat_kv,
    wait_for_compression,
)

__all__ = [
    "AdaKVCluster", "DynamicCacheSplitHeadFlatten", "H2OKVCluster", "HeadKVCluster", "L2NormCluster", "init_headkv", "init_l2norm", "init_think", "PyramidKVCluster", "SnapKVCluster",
    "StreamingLLMKVCluster", "init_adakv", "init_H2O", "init_pyramidkv", "init_snapkv", "init_StreamingLLM",
    "pyramid_budget", "repeat_kv", "wait_for_compression", "replace_llama", "replace_mistral",
]
