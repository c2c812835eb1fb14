"""Pin the oracle: replay it against vectors produced by the unmodified reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the live reference."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import kvc_oracle as O
from tests.helpers import DTYPES, from_bits, golden_cases, load_golden, to_bits

REF = "/root/reference"


def _run_oracle(c, Q, K, V):
    kn = dict(c["knobs"])
    m = c["method"]
    if m == "snapkv":
        return O.snapkv_update_kv(K, Q, V, **kn)
    if m == "pyramidkv":
        return O.pyramidkv_update_kv(K, Q, V, **kn)
    if m == "h2o":
        return O.h2o_update_kv(K, Q, V, **kn)
    if m == "streamingllm":
        return O.streamingllm_update_kv(K, Q, V, **kn)
    raise ValueError(m)


@pytest.mark.parametrize("name", golden_cases("snapkv") + golden_cases("pyramidkv") + golden_cases("h2o")
                         + golden_cases("streamingllm"))
def test_dense_policies_match_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    dtype = DTYPES[c["dtype"]]
    Kc, Vc, idx, pooled = _run_oracle(c, Q, K, V)
    if "pooled" in z:
        # same ATen ops, same order, same machine class: the pooled scores are bit-identical
        assert np.array_equal(to_bits(pooled), z["pooled"]), "pooled scores differ from the reference"
        assert np.array_equal(idx.numpy(), z["idx"]), "topk indices differ from the reference"
    assert Kc.dtype == dtype and tuple(Kc.shape) == tuple(z["K_out"].shape)
    assert np.array_equal(to_bits(Kc), z["K_out"])
    assert np.array_equal(to_bits(Vc), z["V_out"])


@pytest.mark.parametrize("name", golden_cases("l2norm"))
def test_l2norm_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    Kc, Vc, idx, norms = O.l2norm_update_kv(K, Q, V, **c["knobs"])
    assert tuple(Kc.shape) == tuple(z["K_out"].shape)
    assert np.array_equal(to_bits(Kc), z["K_out"]) and np.array_equal(to_bits(Vc), z["V_out"])
    if norms is not None:
        assert np.array_equal(to_bits(norms), z["norms"])


@pytest.mark.parametrize("name", golden_cases("think"))
def test_think_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    res = O.think_update_kv(K, Q, V, **c["knobs"])
    if name == "think_skip":
        assert len(res) == 2 and np.array_equal(to_bits(res[0]), z["K_out"]) and np.array_equal(to_bits(res[1]), z["V_out"])
        return
    pruned, recent, mask, Vc, Kc, key = res
    assert np.array_equal(to_bits(Kc), z["K_out"]) and np.array_equal(to_bits(Vc), z["V_out"])
    assert np.array_equal(to_bits(key), z["think_key"])
    assert np.array_equal(mask.numpy(), z["think_mask"])
    assert tuple(pruned.shape) == tuple(z["think_pruned"].shape)
    assert np.array_equal(to_bits(pruned), z["think_pruned"]) and np.array_equal(to_bits(recent), z["think_recent"])


@pytest.mark.parametrize("name", golden_cases("adakv"))
def test_adakv_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    flatK, flatV, meta, pooled, picks = O.adakv_update_kv(K, Q, V, **c["knobs"])
    assert np.array_equal(to_bits(pooled), z["pooled"])
    for a in ("head_lens", "cu_klen", "cu_headlens", "cu_qlen", "cu_offset", "cu_head_offset", "layer_qlens"):
        assert np.array_equal(meta[a], z["meta_" + a]), a
        assert meta[a].dtype == z["meta_" + a].dtype == np.int32, a
    assert [meta["klen_sum"], meta["qlen_sum"], meta["max_seqlen_k"]] == z["meta_scalars"].tolist()
    assert np.array_equal(to_bits(flatK), z["K_out"])
    assert np.array_equal(to_bits(flatV), z["V_out"])


@pytest.mark.parametrize("name", golden_cases("headkv"))
def test_headkv_matches_reference_vectors(name):
    c, z, (Q, K, V) = load_golden(name)
    kn = dict(c["knobs"])
    flatK, flatV, meta, pooled, picks = O.headkv_update_kv(K, Q, V, kn.pop("head_capacity"), kn.pop("layer_idx"), **kn)
    assert np.array_equal(to_bits(pooled), z["pooled"])
    for a in ("head_lens", "cu_klen", "cu_headlens"):
        assert np.array_equal(meta[a], z["meta_" + a]), a
    assert np.array_equal(to_bits(flatK), z["K_out"]) and np.array_equal(to_bits(flatV), z["V_out"])


def test_pyramid_budget_schedule_matches_survey_numbers():
    # cfg-2 budgets 3932 - 123*l, sum of kept rows 65 840; cfg-5 steps 48.
    b = [O.pyramidkv_budget(32768, 32, 2048, 32, l) for l in range(32)]
    assert b[0] == 3932 and b[31] == 119 and b[0] - b[1] == 123
    assert sum(x + 32 for x in b) == 65840
    b70 = [O.pyramidkv_budget(131072, 32, 2048, 80, l) for l in range(80)]
    assert b70[0] == 3932 and b70[0] - b70[1] == 48 and b70[79] == 140
    assert sum(x + 32 for x in b70) == 165440


def test_flatten_oracle_hand_case():
    # two heads with 2 and 1 rows, d=2: rows keep their order, head h shifts down by h rows and the
    # new row lands after the head's old rows (cuda_api.cu:26-52).
    cache = np.array([[1, 1], [2, 2], [3, 3]], dtype=np.uint16)
    state = np.array([[9, 9], [8, 8]], dtype=np.uint16)
    out = O.update_flatten_view(cache, state, np.array([2, 1], np.int32), np.array([0, 2, 3], np.int32))
    assert out.tolist() == [[1, 1], [2, 2], [9, 9], [3, 3], [8, 8]]
    with pytest.raises(TypeError):
        O.update_flatten_view(cache, state, np.array([2, 1], np.int64), np.array([0, 2, 3], np.int32))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("pooling", ["avgpool", "maxpool"])
def test_oracle_equals_live_reference(dtype, pooling):
    """Random shapes, live: oracle(un-repeated K) == reference(repeat_kv(K)) bit for bit."""
    from kvcache_factory_b200.synth import Shape, make_qkv
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import pyramidkv.pyramidkv_utils as ref
    shape = Shape(B=2, Hq=6, Hkv=3, L=333, d=64)
    Q, K, V = make_qkv(shape, dtype, layer_idx=5, dist="peaked", window=16)
    Kr, Vr = ref.repeat_kv(K, 2), ref.repeat_kv(V, 2)
    with contextlib.redirect_stdout(io.StringIO()):
        r_snap = ref.SnapKVCluster(16, 80, 5, pooling).update_kv(Kr, Q, Vr, None, 2)
        r_pyr = ref.PyramidKVCluster(num_hidden_layers=6, window_size=16, max_capacity_prompt=80, kernel_size=5,
                                     pooling=pooling, layer_idx=3).update_kv(Kr, Q, Vr, None, 2)
        r_h2o = ref.H2OKVCluster(16, 80, 5, pooling).update_kv(Kr, Q, Vr, None, 2)
        r_str = ref.StreamingLLMKVCluster(16, 80, 5, pooling).update_kv(Kr, Q, Vr, None, 2)
    o_snap = O.snapkv_update_kv(K, Q, V, 16, 80, 5, pooling)
    o_pyr = O.pyramidkv_update_kv(K, Q, V, 3, 6, 16, 80, 5, pooling)
    o_h2o = O.h2o_update_kv(K, Q, V, 16, 80)
    o_str = O.streamingllm_update_kv(K, Q, V, 16, 80)
    for r, o in ((r_snap, o_snap), (r_pyr, o_pyr), (r_h2o, o_h2o), (r_str, o_str)):
        assert torch.equal(r[0], o[0]) and torch.equal(r[1], o[1])
    # AdaKV (batch 1)
    Q1, K1, V1 = Q[:1], K[:1], V[:1]
    cl = ref.AdaKVCluster(window_size=16, kernel_size=5, pooling=pooling, max_capacity_prompt=80, floor=0.2,
                          normalize=True)
    with contextlib.redirect_stdout(io.StringIO()):
        rk, rv = cl.update_kv(ref.repeat_kv(K1, 2), Q1, ref.repeat_kv(V1, 2))
    ok, ov, meta, _, _ = O.adakv_update_kv(K1, Q1, V1, 16, 5, pooling, 80, 0.2, True)
    assert torch.equal(rk, ok) and torch.equal(rv, ov)
    assert np.array_equal(cl.cu_klen.numpy(), meta["cu_klen"])
    assert np.array_equal(cl.head_lens.numpy(), meta["head_lens"])
