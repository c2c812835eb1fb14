"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Run from the repo root:   python tests/golden/make_golden.py

The reference (``/root/reference/pyramidkv/pyramidkv_utils.py``) has no tests or fixtures of its
own, so parity is pinned on what the reference itself computes: this script imports
it (read-only, from where it lies), silences its per-call ``print``, runs every cluster named by
SURVEY.md §8(a) on small seeded inputs from ``kvcache_factory_b200.synth`` and writes

    tests/golden/<case>.npz    pooled scores fed to topk (captured by wrapping Tensor.topk),
                               selected indices, compressed K'/V' (bit patterns), AdaKV metadata,
                               and a digest of the inputs so RNG drift is detected, not hidden.

``/root/reference`` does not exist on the GPU box; the committed ``.npz`` files are what travels.
The reference's CUDA-only ``update_flatten_view`` cannot run here (no GPU): its golden vectors come
from ``oracle/flatten_oracle.c``-independent hand-derived cases in ``tests/test_flatten.py`` and,
on the GPU box, from the reference ``.cu`` compiled into ``oracle/_ref`` when present.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from kvcache_factory_b200.synth import Shape, make_qkv, tensor_digest  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> dict(method, shape, dtype, dist, knobs)
CASES = {
    # cfg-1 geometry shrunk in L only (same heads/d/window/pool) so the file stays small
    "snapkv_bf16_L512":   dict(method="snapkv", B=1, Hq=8, Hkv=2, L=512, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(window_size=32, max_capacity_prompt=128, kernel_size=7, pooling="avgpool")),
    "snapkv_fp16_max":    dict(method="snapkv", B=2, Hq=4, Hkv=4, L=300, d=128, dtype="fp16", dist="peaked",
                               knobs=dict(window_size=8, max_capacity_prompt=64, kernel_size=5, pooling="maxpool")),
    "snapkv_bf16_normal": dict(method="snapkv", B=1, Hq=4, Hkv=1, L=640, d=128, dtype="bf16", dist="normal",
                               knobs=dict(window_size=32, max_capacity_prompt=96, kernel_size=7, pooling="avgpool")),
    "snapkv_passthrough": dict(method="snapkv", B=1, Hq=4, Hkv=2, L=100, d=128, dtype="bf16", dist="normal",
                               knobs=dict(window_size=32, max_capacity_prompt=128, kernel_size=5, pooling="avgpool")),
    "pyramidkv_l0":       dict(method="pyramidkv", B=1, Hq=8, Hkv=2, L=1024, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(num_hidden_layers=8, layer_idx=0, window_size=32, max_capacity_prompt=160,
                                          kernel_size=7, pooling="avgpool")),
    "pyramidkv_l7":       dict(method="pyramidkv", B=1, Hq=8, Hkv=2, L=1024, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(num_hidden_layers=8, layer_idx=7, window_size=32, max_capacity_prompt=160,
                                          kernel_size=7, pooling="avgpool")),
    "pyramidkv_mid":      dict(method="pyramidkv", B=1, Hq=4, Hkv=2, L=200, d=128, dtype="fp16", dist="peaked",
                               knobs=dict(num_hidden_layers=4, layer_idx=2, window_size=16, max_capacity_prompt=128,
                                          kernel_size=5, pooling="maxpool")),   # q_len < 2*(cap-w) branch
    "h2o_bf16":           dict(method="h2o", B=1, Hq=4, Hkv=2, L=384, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(window_size=32, max_capacity_prompt=96)),
    "streamingllm":       dict(method="streamingllm", B=2, Hq=4, Hkv=2, L=256, d=128, dtype="bf16", dist="normal",
                               knobs=dict(window_size=16, max_capacity_prompt=64)),
    "adakv_bf16":         dict(method="adakv", B=1, Hq=8, Hkv=2, L=768, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=128,
                                          floor=0.2, normalize=True)),
    "adakv_fp16_nonorm":  dict(method="adakv", B=1, Hq=4, Hkv=4, L=512, d=128, dtype="fp16", dist="peaked",
                               knobs=dict(window_size=16, kernel_size=5, pooling="avgpool", max_capacity_prompt=96,
                                          floor=0.5, normalize=False)),
    "headkv_bf16":        dict(method="headkv", B=1, Hq=8, Hkv=2, L=768, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=128,
                                          layer_idx=1, head_capacity=[[96] * 8, [10, 200, 96, 1, 736, 64, 300, 50]])),
    "l2norm_bf16":        dict(method="l2norm", B=1, Hq=8, Hkv=2, L=640, d=128, dtype="bf16", dist="normal",
                               knobs=dict(max_capacity_prompt=96, layer_idx=3, skip_layers=[0, 1])),
    "l2norm_fp16":        dict(method="l2norm", B=2, Hq=4, Hkv=4, L=300, d=128, dtype="fp16", dist="peaked",
                               knobs=dict(max_capacity_prompt=64, layer_idx=2, skip_layers=[])),
    "l2norm_skip":        dict(method="l2norm", B=1, Hq=4, Hkv=2, L=200, d=128, dtype="bf16", dist="normal",
                               knobs=dict(max_capacity_prompt=64, layer_idx=1, skip_layers=[0, 1])),
    "think_bf16":         dict(method="think", B=1, Hq=8, Hkv=2, L=700, d=128, dtype="bf16", dist="peaked",
                               knobs=dict(window_size=32, max_capacity_prompt=160, kernel_size=5, pooling="avgpool",
                                          recent_size=32, ratio=0.4)),
    "think_fp16":         dict(method="think", B=1, Hq=4, Hkv=4, L=500, d=128, dtype="fp16", dist="normal",
                               knobs=dict(window_size=16, max_capacity_prompt=128, kernel_size=7, pooling="maxpool",
                                          recent_size=16, ratio=0.25)),
    "think_skip":         dict(method="think", B=1, Hq=4, Hkv=2, L=100, d=128, dtype="bf16", dist="normal",
                               knobs=dict(window_size=32, max_capacity_prompt=160, kernel_size=5, pooling="avgpool",
                                          recent_size=32, ratio=0.4)),
    "adakv_nocompress":   dict(method="adakv", B=1, Hq=4, Hkv=2, L=96, d=128, dtype="bf16", dist="normal",
                               knobs=dict(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=128,
                                          floor=0.2, normalize=True)),
}

DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous().cpu()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


@contextlib.contextmanager
def capture_topk(store: list):
    """Record the tensor each ``Tensor.topk`` call is applied to (the pooled scores)."""
    orig = torch.Tensor.topk

    def spy(self, *a, **kw):
        out = orig(self, *a, **kw)
        store.append((self.detach().clone(), out.indices.detach().clone()))
        return out

    torch.Tensor.topk = spy
    orig_fn = torch.topk

    def spy_fn(x, *a, **kw):
        out = orig_fn(x, *a, **kw)
        store.append((x.detach().clone(), out.indices.detach().clone()))
        return out

    torch.topk = spy_fn
    try:
        yield
    finally:
        torch.Tensor.topk = orig
        torch.topk = orig_fn


def run_case(ref, name: str, c: dict) -> dict:
    shape = Shape(c["B"], c["Hq"], c["Hkv"], c["L"], c["d"])
    dtype = DTYPES[c["dtype"]]
    kn = dict(c["knobs"])
    w = kn.get("window_size", 32)
    Q, K, V = make_qkv(shape, dtype, layer_idx=kn.get("layer_idx", 0), dist=c["dist"], window=w)
    Kr, Vr = ref.repeat_kv(K, shape.groups), ref.repeat_kv(V, shape.groups)
    out = {"inputs_digest": np.array(tensor_digest(Q, K, V))}
    calls: list = []
    with contextlib.redirect_stdout(io.StringIO()), capture_topk(calls):
        m = c["method"]
        if m == "snapkv":
            cl = ref.SnapKVCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr, None, shape.groups)
        elif m == "pyramidkv":
            cl = ref.PyramidKVCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr, None, shape.groups)
        elif m == "h2o":
            cl = ref.H2OKVCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr, None, shape.groups)
        elif m == "streamingllm":
            cl = ref.StreamingLLMKVCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr, None, shape.groups)
        elif m == "think":
            cl = ref.SnapKVCluster(**kn)
            res = cl.update_think(Kr, Q, Vr, None, shape.groups)
            if len(res) == 2:                                         # below the capacity: pass-through (:357-358)
                Kc, Vc = res
            else:
                pruned, recent, mask, Vc = res
                out["think_pruned"] = bits(pruned)
                out["think_recent"] = bits(recent)
                out["think_mask"] = mask.numpy()
                # K' itself is not returned by update_think: the same cluster's update_kv gives it
                Kc, _ = ref.SnapKVCluster(**kn).update_kv(Kr, Q, Vr, None, shape.groups)
                out["think_key"] = bits(calls[1][0])                  # the channel scores topk(largest=False) ranks
        elif m == "l2norm":
            cl = ref.L2NormCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr, None, shape.groups)
            out["norms"] = bits(torch.norm(Kr, p=2, dim=-1))          # what update_kv sorts (:420)
        elif m in ("adakv", "headkv"):
            cl = ref.AdaKVCluster(**kn) if m == "adakv" else ref.HeadKVCluster(**kn)
            Kc, Vc = cl.update_kv(Kr, Q, Vr)
            for a in ("head_lens", "cu_klen", "cu_headlens", "cu_qlen", "cu_offset", "cu_head_offset",
                      "layer_qlens"):
                out["meta_" + a] = getattr(cl, a).cpu().numpy()
            out["meta_scalars"] = np.array([cl.klen_sum, cl.qlen_sum, cl.max_seqlen_k], dtype=np.int64)
            out["pooled"] = bits(cl.calcul_attn_sore(Kr, Q))
        else:
            raise ValueError(m)
    out["K_out"] = bits(Kc)
    out["V_out"] = bits(Vc)
    if m in ("snapkv", "pyramidkv", "h2o", "think") and calls:
        pooled, idx = calls[0]
        out["pooled"] = bits(pooled)
        out["idx"] = idx.numpy().astype(np.int64)
    if m == "adakv" and calls:
        out["global_topk_input"] = bits(calls[0][0])
    return out


def main() -> None:
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import pyramidkv.pyramidkv_utils as ref  # the unmodified reference
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    manifest = {}
    only = set(sys.argv[1:])
    if only:   # add / refresh selected cases only, keep the rest of the manifest
        with open(os.path.join(OUT, "manifest.json")) as f:
            manifest = json.load(f)
    for name, c in CASES.items():
        if only and name not in only:
            continue
        res = run_case(ref, name, c)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
        manifest[name] = {k: v for k, v in c.items()}
        manifest[name]["torch"] = torch.__version__
        size = os.path.getsize(os.path.join(OUT, name + ".npz"))
        print(f"{name:24s} K_out {tuple(res['K_out'].shape)}  {size/1024:.1f} KiB")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
