"""Synthetic prefill workloads for the KV-compression hot path.

The generators follow SURVEY.md §8(d): per layer a CPU ``torch.Generator`` seeded with
``1234 + layer_idx`` draws Q ``[B,Hq,L,d]`` and un-repeated K, V ``[B,Hkv,L,d]`` from N(0,1) in fp32
and casts to the model dtype.  The "peaked" distribution additionally plants heavy-hitter keys
(``K[b,h,j,:] += 4 * qbar[b,h]`` for 1 % random ``j``) so that top-k is meaningful and score ties
at the selection threshold are rare.

Everything here is host-side test/bench plumbing; nothing on the product path depends on it.
"""
from __future__ import annotations

import dataclasses
import hashlib

import torch

SEED_BASE = 1234


@dataclasses.dataclass(frozen=True)
class Shape:
    """Head geometry of one prefill ``update_kv`` call."""

    B: int
    Hq: int
    Hkv: int
    L: int
    d: int = 128

    @property
    def groups(self) -> int:
        return self.Hq // self.Hkv


# Model head geometries named by BASELINE.json (SURVEY.md §8 notation).
LLAMA3_8B = dict(Hq=32, Hkv=8, d=128, layers=32)
MISTRAL_7B = dict(Hq=32, Hkv=8, d=128, layers=32)
LLAMA3_70B = dict(Hq=64, Hkv=8, d=128, layers=80)


def make_qkv(shape: Shape, dtype: torch.dtype = torch.bfloat16, layer_idx: int = 0,
             dist: str = "normal", window: int = 32, seed_base: int = SEED_BASE,
             q_window_only: bool = False):
    """Return (Q, K, V) CPU tensors for one layer.

    ``q_window_only`` draws only the last ``window`` query rows (the only rows the window-scored
    policies read) and leaves the rest zero; the stream of random numbers for K and V is the same
    either way because K and V are drawn first.
    """
    g = torch.Generator().manual_seed(seed_base + layer_idx)
    B, Hq, Hkv, L, d = shape.B, shape.Hq, shape.Hkv, shape.L, shape.d
    K = torch.randn(B, Hkv, L, d, generator=g, dtype=torch.float32)
    V = torch.randn(B, Hkv, L, d, generator=g, dtype=torch.float32)
    if q_window_only:
        Q = torch.zeros(B, Hq, L, d, dtype=torch.float32)
        Q[:, :, L - window:, :] = torch.randn(B, Hq, window, d, generator=g, dtype=torch.float32)
    else:
        Q = torch.randn(B, Hq, L, d, generator=g, dtype=torch.float32)
    if dist == "peaked":
        n_hot = max(1, (L - window) // 100)
        grp = shape.groups
        for b in range(B):
            for h in range(Hkv):
                qbar = Q[b, h * grp, L - window:, :].mean(dim=0)
                hot = torch.randperm(L - window, generator=g)[:n_hot]
                K[b, h, hot, :] += 4.0 * qbar
    elif dist != "normal":
        raise ValueError(f"unknown distribution {dist!r}")
    return Q.to(dtype), K.to(dtype), V.to(dtype)


def tensor_digest(*tensors: torch.Tensor) -> str:
    """Stable hex digest of the raw bytes of CPU tensors (pins the RNG stream used by fixtures)."""
    h = hashlib.sha256()
    for t in tensors:
        t = t.detach().contiguous().cpu()
        h.update(t.view(torch.uint8).numpy().tobytes())
    return h.hexdigest()[:32]
