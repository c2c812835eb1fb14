"""Shared helpers for the parity tests (golden loading, dtype plumbing, ulp distances)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from kvcache_factory_b200.synth import Shape, make_qkv, tensor_digest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}

with open(os.path.join(GOLDEN_DIR, "manifest.json")) as _f:
    MANIFEST = json.load(_f)


def golden_cases(method=None):
    return sorted(n for n, c in MANIFEST.items() if method is None or c["method"] == method)


def load_golden(name):
    """Return (case dict, npz dict, (Q, K, V) regenerated from the seed and digest-checked)."""
    c = MANIFEST[name]
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    shape = Shape(c["B"], c["Hq"], c["Hkv"], c["L"], c["d"])
    kn = c["knobs"]
    Q, K, V = make_qkv(shape, DTYPES[c["dtype"]], layer_idx=kn.get("layer_idx", 0), dist=c["dist"],
                       window=kn.get("window_size", 32))
    assert tensor_digest(Q, K, V) == str(z["inputs_digest"]), (
        "synthetic input stream drifted from the one the golden vectors were made with")
    return c, z, (Q, K, V)


def from_bits(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    if dtype in (torch.bfloat16, torch.float16):
        return torch.from_numpy(a.view(np.int16).copy()).view(dtype)
    return torch.from_numpy(a.copy())


def to_bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous().cpu()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def ulp_of(dtype: torch.dtype) -> float:
    """Relative spacing of the model dtype (2^-7 for bf16, 2^-10 for fp16, 2^-23 for fp32)."""
    return {torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10, torch.float32: 2.0 ** -23}[dtype]


def score_mismatch(test: torch.Tensor, ref: torch.Tensor, dtype: torch.dtype, ulps: float):
    """Fraction of elements whose |test-ref| exceeds ``ulps`` model-dtype ulps of |ref|, and the
    worst ratio; scores are compared in fp32."""
    t, r = test.float().cpu(), ref.float().cpu()
    tol = ulps * ulp_of(dtype) * r.abs().clamp_min(torch.finfo(dtype).tiny)
    err = (t - r).abs()
    bad = err > tol
    worst = float((err / tol).max()) if err.numel() else 0.0
    return float(bad.float().mean()), worst
