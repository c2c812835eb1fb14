"""ctypes binding of ``libkvc_b200.so`` (C ABI declared in ``include/kvc_b200.h``).

PyTorch is used here for device memory, streams and nothing else: every function marshals
``tensor.data_ptr()``, element strides and the current CUDA stream into the C ABI.  There is no CPU
path and no alternative backend: if the library is missing or the tensors are not on a CUDA device
the call raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# KVC_LIB_PATH: developer override (e.g. the `make TRACE=1` build, libkvc_b200_trace.so)
LIB_PATH = os.environ.get("KVC_LIB_PATH") or os.path.join(_PKG_DIR, "libkvc_b200.so")

KVC_BF16, KVC_FP16 = 0, 1
POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2
REDUCE_SUM, REDUCE_MEAN = 0, 1
IMPL_AUTO, IMPL_MMA_SYNC, IMPL_TCGEN05, IMPL_TCGEN05_FUSED, IMPL_TCGEN05_HALF, IMPL_TCGEN05_DYN = 0, 1, 2, 3, 4, 5
ABI_VERSION = 1

_POOL = {None: POOL_NONE, "none": POOL_NONE, "avgpool": POOL_AVG, "maxpool": POOL_MAX}

# name -> (restype, argtypes); mirrors include/kvc_b200.h one to one
_i64, _p, _i = c_int64, c_void_p, c_int
_SIGNATURES = {
    "kvc_version": (c_int, []),
    "kvc_last_error": (c_char_p, []),
    "kvc_launch_count": (ctypes.c_uint64, []),
    "kvc_sm_count": (c_int, [POINTER(c_int)]),
    "kvc_window_scores_workspace_bytes": (c_size_t, [_i] * 6),
    "kvc_window_scores": (c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _p,
                                  _p, c_size_t, _i, _p]),
    "kvc_window_scores_lse": (c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p,
                                      _p, c_size_t, _p]),
    "kvc_h2o_scores_workspace_bytes": (c_size_t, [_i] * 6),
    "kvc_h2o_scores": (c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i, _i, _p, _p,
                               c_size_t, _p]),
    "kvc_pool_topk_workspace_bytes": (c_size_t, [_i] * 3),
    "kvc_pool_topk": (c_int, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i64, _p, _p, _p, c_size_t, _p]),
    "kvc_gather_dense": (c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p,
                                 _p, _p]),
    "kvc_key_norm_scores": (c_int, [_p, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i, _p, _p]),
    "kvc_compress_dense_workspace_bytes": (c_size_t, [_i] * 7),
    "kvc_compress_dense": (c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64,
                                   _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, c_size_t,
                                   POINTER(c_size_t), POINTER(c_size_t), _p]),
    "kvc_gather_varlen": (c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p,
                                  _p, _p]),
    "kvc_adakv_allocate_workspace_bytes": (c_size_t, [_i] * 2),
    "kvc_adakv_allocate": (c_int, [_p, _p, _i, _i, _i, _i, c_float, _i, _i, _i, _p, _p, _p, _p, c_size_t, _p]),
    "kvc_adakv_key_histogram": (c_int, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "kvc_adakv_hist_coarse": (c_int, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "kvc_adakv_hist_fine": (c_int, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "kvc_adakv_prefill_workspace_bytes": (c_size_t, [_i] * 6),
    "kvc_adakv_prefill": (c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                  c_float, _i, _i, _i, _p, _p, _i64, _p, _p, _p, _p, _i, _p, c_size_t,
                                  POINTER(c_int), POINTER(c_int64), POINTER(c_size_t), _p]),
    "kvc_think_prune_workspace_bytes": (c_size_t, [_i] * 4),
    "kvc_think_prune": (c_int, [_p, _p, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, c_size_t, _p]),
    "kvc_update_flatten_view": (c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i64, _p]),
    "kvc_append_rows": (c_int, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "kvc_flatten_regrow": (c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i64, _p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class KvcError(RuntimeError):
    """Non-zero status from libkvc_b200 (negative: argument/capability, positive: cudaError_t)."""

    def __init__(self, fn: str, status: int, message: str):
        super().__init__(f"{fn} failed with status {status}: {message}")
        self.status = status


def lib() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C kvcache_factory_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. kvcache_factory_b200 has no CPU or "
                "PyTorch fallback for the compression path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if handle.kvc_version() != ABI_VERSION:
            raise RuntimeError(f"libkvc_b200 ABI {handle.kvc_version()} != binding ABI {ABI_VERSION}")
        _lib = handle
    return _lib


def _check(fn: str, status: int) -> None:
    if status != 0:
        msg = lib().kvc_last_error()
        raise KvcError(fn, status, msg.decode() if msg else "")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return KVC_BF16
    if dtype == torch.float16:
        return KVC_FP16
    raise TypeError(f"kvcache_factory_b200 kernels take bf16 or fp16 tensors, got {dtype} "
                    "(the reference's native kernel has the same restriction, csrc/static_switch.h:3-12)")


def pool_code(pooling) -> int:
    if pooling not in _POOL:
        raise ValueError("Pooling method not supported")  # pyramidkv_utils.py:237,333
    return _POOL[pooling]


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError("kvcache_factory_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on "
                               f"{t.device}")
        if t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


class _on_device:
    """``with torch.cuda.device(dev)`` only when ``dev`` is not the current device already: the guard costs ~5 us of host
    time per call, which is what the eager drop-in path of short prompts is bound by."""
    __slots__ = ("dev", "guard")

    def __init__(self, dev: torch.device):
        self.dev, self.guard = dev, None

    def __enter__(self):
        if self.dev.index is not None and torch.cuda.current_device() != self.dev.index:
            self.guard = torch.cuda.device(self.dev)
            self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(dev: torch.device) -> int:
    """Handle of the caller's current stream on ``dev`` (the raw getter is ~7 us cheaper than building a Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(dev.index if dev.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(dev).cuda_stream


def _rowmajor_last(t: torch.Tensor, align_elems: int) -> torch.Tensor:
    """Kernels need a unit-stride last dim and (for K/V) 16-byte aligned rows; HF tensors have both.
    Anything else is repacked once on the device (a copy kernel, not a compute fallback)."""
    st = t.stride()
    ok = st[-1] == 1 and t.data_ptr() % (2 * align_elems) == 0
    if ok:
        for s in st[:-1]:
            if s % align_elems:
                ok = False
                break
    return t if ok else t.contiguous()


# ------------------------------------------------------------------------------------------------
# op wrappers
# ------------------------------------------------------------------------------------------------
def window_scores(query_states: torch.Tensor, key_states: torch.Tensor, window_size: int,
                  reduce: str = "sum", impl: int = IMPL_AUTO, seq_len: int = None,
                  lse: torch.Tensor = None) -> torch.Tensor:
    """[B,Hq,L-w] scores of every non-window key under the last ``window_size`` queries.

    ``query_states`` is ``[B,Hq,L,d]`` or, with ``seq_len=L``, only its last ``window_size`` rows
    ``[B,Hq,w,d]`` (the C ABI addresses row L-w+i of the logical tensor, so the base pointer is
    simply moved back by L-w rows; nothing outside the window is ever dereferenced).
    """
    dev = _require_cuda(query_states, key_states)
    B, Hq, Lq, d = query_states.shape
    Hkv, L = key_states.shape[1], key_states.shape[2]
    code = dtype_code(key_states.dtype)
    if query_states.dtype != key_states.dtype:
        raise TypeError("query and key dtypes differ")
    q = _rowmajor_last(query_states, 2)
    k = _rowmajor_last(key_states, 8)
    q_ptr = q.data_ptr()
    if Lq != L:
        if not (seq_len == L and Lq == window_size):
            raise ValueError(f"query_states has {Lq} rows, key_states {L}")
        q_ptr -= (L - window_size) * q.stride(2) * q.element_size()
    out = torch.empty((B, Hq, L - window_size), dtype=key_states.dtype, device=dev)
    if lse is not None:
        # statistics supplied (float32 [B,Hq,w], e.g. flash attention's softmax_lse of the window rows): pass 2 only
        _require_cuda(key_states, lse)
        if lse.dtype != torch.float32 or tuple(lse.shape) != (B, Hq, window_size):
            raise TypeError(f"lse must be float32 [B,Hq,w] = {(B, Hq, window_size)}, got {lse.dtype} {tuple(lse.shape)}")
        lse = lse.contiguous()
        q8 = _rowmajor_last(query_states, 8)
        q_ptr = q8.data_ptr()
        if Lq != L:
            q_ptr -= (L - window_size) * q8.stride(2) * q8.element_size()
        with _on_device(dev):
            ws_bytes = lib().kvc_window_scores_workspace_bytes(B, Hq, Hkv, L, d, window_size)
            ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
            _check("kvc_window_scores_lse", lib().kvc_window_scores_lse(
                q_ptr, q8.stride(0), q8.stride(1), q8.stride(2), k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
                lse.data_ptr(), B, Hq, Hkv, L, d, window_size, code, REDUCE_MEAN if reduce == "mean" else REDUCE_SUM,
                out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
        return out
    with _on_device(dev):
        ws_bytes = lib().kvc_window_scores_workspace_bytes(B, Hq, Hkv, L, d, window_size)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        _check("kvc_window_scores", lib().kvc_window_scores(
            q_ptr, q.stride(0), q.stride(1), q.stride(2),
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            B, Hq, Hkv, L, d, window_size, code, REDUCE_MEAN if reduce == "mean" else REDUCE_SUM,
            out.data_ptr(), ws.data_ptr(), ws.numel(), impl, _stream(dev)))
    return out


def h2o_scores(query_states: torch.Tensor, key_states: torch.Tensor, window_size: int) -> torch.Tensor:
    dev = _require_cuda(query_states, key_states)
    B, Hq, L, d = query_states.shape
    Hkv = key_states.shape[1]
    code = dtype_code(key_states.dtype)
    q = _rowmajor_last(query_states, 8)
    k = _rowmajor_last(key_states, 8)
    out = torch.empty((B, Hq, L - window_size), dtype=key_states.dtype, device=dev)
    with _on_device(dev):
        ws_bytes = lib().kvc_h2o_scores_workspace_bytes(B, Hq, Hkv, L, d, window_size)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        _check("kvc_h2o_scores", lib().kvc_h2o_scores(
            q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            B, Hq, Hkv, L, d, window_size, code, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return out


def pool_topk(scores: torch.Tensor, kernel_size: int, pooling, k: int, k_per_row: torch.Tensor = None,
              want_pooled: bool = False, want_sums: bool = False, sums_only: bool = False):
    """Pool + per-row top-k.  Returns (idx int32 [..., k], pooled or None, sums float[rows,2] or None).

    ``sums_only``: only the pooled scores and the (top-k, all) sums are wanted; no indices are produced
    (idx is returned as None) and the kernel stops after the threshold search."""
    if sums_only:
        try:
            return _pool_topk(scores, kernel_size, pooling, k, k_per_row, want_pooled, True, no_idx=True)
        except KvcError as e:
            if e.status != -2:
                raise
    return _pool_topk(scores, kernel_size, pooling, k, k_per_row, want_pooled, want_sums or sums_only)


def _pool_topk(scores, kernel_size, pooling, k, k_per_row, want_pooled, want_sums, no_idx=False):
    dev = _require_cuda(scores)
    code = dtype_code(scores.dtype)
    pcode = pool_code(pooling)
    lead, n = scores.shape[:-1], scores.shape[-1]
    rows = 1
    for s in lead:
        rows *= s
    scores = scores.contiguous()
    idx = None if no_idx else torch.empty((*lead, k), dtype=torch.int32, device=dev)
    pooled = torch.empty_like(scores) if want_pooled else None
    sums = torch.empty((rows, 2), dtype=torch.float32, device=dev) if want_sums else None
    if k_per_row is not None:
        if k_per_row.dtype != torch.int32 or k_per_row.numel() != rows:
            raise TypeError("k_per_row must be an int32 tensor with one entry per row")
        _require_cuda(scores, k_per_row)
    with _on_device(dev):
        ws_bytes = lib().kvc_pool_topk_workspace_bytes(rows, n, k)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        _check("kvc_pool_topk", lib().kvc_pool_topk(
            scores.data_ptr(), rows, n, code, kernel_size, pcode, k,
            k_per_row.data_ptr() if k_per_row is not None else None, idx.data_ptr() if idx is not None else None, k,
            pooled.data_ptr() if pooled is not None else None, sums.data_ptr() if sums is not None else None,
            ws.data_ptr(), ws.numel(), _stream(dev)))
    return idx, pooled, sums


def gather_dense(key_states: torch.Tensor, value_states: torch.Tensor, idx, num_heads: int, n_sel: int,
                 window_size: int):
    """Compressed (K', V') [B,Hq,n_sel+w,d]; ``idx=None`` keeps the first n_sel rows (StreamingLLM).

    ``value_states`` may instead be a PINNED host tensor (same shape/dtype): V is only ever read at the
    selected rows, so a caller holding its KV in host memory does not have to move the other rows to
    the device -- the kernel's 16-byte loads then read the selected rows straight over PCIe through
    the unified address space (pinned allocations are device-accessible at their host address).
    """
    if value_states.device.type == "cpu":
        dev = _require_cuda(key_states)                              # K is always read on the device
        if not value_states.is_pinned():
            raise RuntimeError("a host value_states tensor must be pinned (page-locked) memory")
    else:
        dev = _require_cuda(key_states, value_states)
    B, Hkv, L, d = key_states.shape
    if value_states.shape != key_states.shape or value_states.dtype != key_states.dtype:
        raise ValueError("key_states and value_states must have the same shape and dtype")
    e = key_states.element_size()
    k = _rowmajor_last(key_states, 16 // e)
    v = _rowmajor_last(value_states, 16 // e)
    if v is not value_states and value_states.device.type == "cpu":
        # a repacked copy would be pageable memory whose host address the kernel cannot read
        raise RuntimeError("a pinned host value_states tensor must have a unit-stride last dim and 16-byte aligned rows")
    if idx is not None:
        _require_cuda(key_states, idx)
        if idx.dtype != torch.int32 or not idx.is_contiguous():
            raise TypeError("idx must be a contiguous int32 tensor")
    out_k = torch.empty((B, num_heads, n_sel + window_size, d), dtype=k.dtype, device=dev)
    out_v = torch.empty_like(out_k)
    with _on_device(dev):
        _check("kvc_gather_dense", lib().kvc_gather_dense(
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
            idx.data_ptr() if idx is not None else None, B, num_heads, Hkv, L, d, n_sel, window_size, e,
            out_k.data_ptr(), out_v.data_ptr(), _stream(dev)))
    return out_k, out_v


def key_norm_scores(key_states: torch.Tensor, num_heads: int) -> torch.Tensor:
    """[B,Hq,L] NEGATED key L2 norms in the model dtype (L2Norm policy: top-k of this = smallest norms, ascending)."""
    dev = _require_cuda(key_states)
    B, Hkv, L, d = key_states.shape
    code = dtype_code(key_states.dtype)
    k = _rowmajor_last(key_states, 8)
    out = torch.empty((B, num_heads, L), dtype=key_states.dtype, device=dev)
    with _on_device(dev):
        _check("kvc_key_norm_scores", lib().kvc_key_norm_scores(
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), B, num_heads, Hkv, L, d, code, out.data_ptr(),
            _stream(dev)))
    return out


_ws_bytes_cache = {}


def compress_dense(query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor, window_size: int,
                   kernel_size: int, pooling, n_keep: int, num_heads: int, impl: int = IMPL_AUTO, seq_len: int = None,
                   want_debug: bool = False):
    """The whole dense update_kv in ONE library call: (K', V') [B,Hq,n_keep+w,d] (+ scores / indices views of the
    workspace when ``want_debug``).  ``pooling == "streaming"`` keeps the first n_keep rows without scoring."""
    dev = _require_cuda(key_states, value_states) if query_states is None else _require_cuda(query_states, key_states, value_states)
    B, Hkv, L, d = key_states.shape
    if value_states.shape != key_states.shape or value_states.dtype != key_states.dtype:
        raise ValueError("key_states and value_states must have the same shape and dtype")
    code = dtype_code(key_states.dtype)
    e = key_states.element_size()
    k = _rowmajor_last(key_states, 16 // e)
    v = _rowmajor_last(value_states, 16 // e)
    streaming = pooling == "streaming"
    if streaming:
        q, q_ptr, qs, pcode, ksz = None, None, (0, 0, 0), POOL_NONE, 0
    else:
        if query_states.dtype != key_states.dtype:
            raise TypeError("query and key dtypes differ")
        pcode, ksz = pool_code(pooling), kernel_size
        q = _rowmajor_last(query_states, 8 if d == 128 else 2)
        q_ptr, Lq = q.data_ptr(), q.shape[2]
        if Lq != L:
            if not (seq_len == L and Lq == window_size):
                raise ValueError(f"query_states has {Lq} rows, key_states {L}")
            q_ptr -= (L - window_size) * q.stride(2) * q.element_size()
        qs = (q.stride(0), q.stride(1), q.stride(2))
    out_k = torch.empty((B, num_heads, n_keep + window_size, d), dtype=k.dtype, device=dev)
    out_v = torch.empty_like(out_k)
    key = (dev.index, B, num_heads, Hkv, L, d, window_size, n_keep)
    with _on_device(dev):
        nbytes = _ws_bytes_cache.get(key)
        if nbytes is None:
            nbytes = _ws_bytes_cache[key] = max(256, lib().kvc_compress_dense_workspace_bytes(B, num_heads, Hkv, L, d, window_size, n_keep))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        io, so = c_size_t(0), c_size_t(0)
        _check("kvc_compress_dense", lib().kvc_compress_dense(
            q_ptr, qs[0], qs[1], qs[2], k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), B, num_heads, Hkv, L, d, window_size, code,
            ksz, pcode, n_keep, impl, out_k.data_ptr(), out_v.data_ptr(), ws.data_ptr(), nbytes,
            ctypes.byref(io), ctypes.byref(so), _stream(dev)))
    if not want_debug or streaming:
        return out_k, out_v, None, None
    n = L - window_size
    scores = ws[so.value: so.value + B * num_heads * n * e].view(key_states.dtype).view(B, num_heads, n)
    idx = ws[io.value: io.value + B * num_heads * n_keep * 4].view(torch.int32).view(B, num_heads, n_keep)
    return out_k, out_v, scores, idx


def gather_varlen(key_states, value_states, idx, cap, cu_klen, num_heads: int, window_size: int, max_cap: int,
                  total_rows: int):
    """Flattened (K', V') [total_rows, d]: head h owns rows [cu_klen[h], cu_klen[h+1])."""
    dev = _require_cuda(key_states, value_states, idx, cap, cu_klen)
    B, Hkv, L, d = key_states.shape
    assert B == 1                                                   # pyramidkv_utils.py:724
    e = key_states.element_size()
    k = _rowmajor_last(key_states, 16 // e)
    v = _rowmajor_last(value_states, 16 // e)
    out_k = torch.empty((total_rows, d), dtype=k.dtype, device=dev)
    out_v = torch.empty_like(out_k)
    with _on_device(dev):
        _check("kvc_gather_varlen", lib().kvc_gather_varlen(
            k.data_ptr(), k.stride(1), k.stride(2), v.data_ptr(), v.stride(1), v.stride(2),
            idx.data_ptr(), idx.stride(-2), cap.data_ptr(), cu_klen.data_ptr(), num_heads, Hkv, L, d, window_size,
            max_cap, e, out_k.data_ptr(), out_v.data_ptr(), _stream(dev)))
    return out_k, out_v


def adakv_allocate(pooled: torch.Tensor, sums: torch.Tensor, base_capacity: int, floor_ratio: float,
                   normalize: bool, window_size: int):
    """Device int32 tensors (cap[H], head_lens[H], cu_klen[H+1]) from pooled [H,n] scores."""
    dev = _require_cuda(pooled, sums)
    H, n = pooled.shape
    code = dtype_code(pooled.dtype)
    cap = torch.empty(H, dtype=torch.int32, device=dev)
    head_lens = torch.empty(H, dtype=torch.int32, device=dev)
    cu_klen = torch.empty(H + 1, dtype=torch.int32, device=dev)
    # torch evaluates `counts * (1 - floor_ratio)` with the Python double rounded to fp32
    one_minus_floor = float(torch.tensor(1 - floor_ratio, dtype=torch.float32))
    with _on_device(dev):
        ws_bytes = lib().kvc_adakv_allocate_workspace_bytes(H, n)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        _check("kvc_adakv_allocate", lib().kvc_adakv_allocate(
            pooled.data_ptr(), sums.data_ptr(), H, n, code, base_capacity, one_minus_floor,
            int(base_capacity * floor_ratio), int(bool(normalize)), window_size,
            cap.data_ptr(), head_lens.data_ptr(), cu_klen.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return cap, head_lens, cu_klen


_pinned_caps = {}


def adakv_prefill(query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor, window_size: int,
                  kernel_size: int, pooling, base_capacity: int, floor_ratio: float, normalize: bool,
                  impl: int = IMPL_AUTO, want_debug: bool = False):
    """``AdaKVCluster.update_kv`` in ONE library call (scores, allocation, capacity read, select, varlen gather).
    Returns (flat_k, flat_v, cap, head_lens, cu_klen, max_cap, klen_sum, debug)."""
    dev = _require_cuda(query_states, key_states, value_states)
    B, Hkv, L, d = key_states.shape
    H = query_states.shape[1]
    if B != 1:
        raise AssertionError("AdaKV's flattened cache is defined for batch 1")      # pyramidkv_utils.py:724
    if query_states.dtype != key_states.dtype or value_states.dtype != key_states.dtype:
        raise TypeError("query, key and value dtypes differ")
    e = key_states.element_size()
    k = _rowmajor_last(key_states, 16 // e)
    v = _rowmajor_last(value_states, 16 // e)
    q = _rowmajor_last(query_states, 8 if d == 128 else 2)
    w, n = window_size, L - window_size
    out_rows = H * (base_capacity + w) + H            # sum of round(count*(1-f) + int(base*f)) <= H*base + H/2
    flat_k = torch.empty((out_rows, d), dtype=k.dtype, device=dev)
    flat_v = torch.empty_like(flat_k)
    cap = torch.empty(H, dtype=torch.int32, device=dev)
    head_lens = torch.empty(H, dtype=torch.int32, device=dev)
    cu_klen = torch.empty(H + 1, dtype=torch.int32, device=dev)
    host = _pinned_caps.get((dev.index, H))
    if host is None:
        host = _pinned_caps[(dev.index, H)] = torch.empty(H, dtype=torch.int32).pin_memory()
    one_minus_floor = float(np.float32(1 - floor_ratio))                          # as torch evaluates it (:719)
    bound = n    # index rows and select scratch sized for the widest possible head: workspace bytes, not work
    with _on_device(dev):
        key = ("adakv", dev.index, H, Hkv, L, d, w, bound)
        nbytes = _ws_bytes_cache.get(key)
        if nbytes is None:
            nbytes = _ws_bytes_cache[key] = max(256, lib().kvc_adakv_prefill_workspace_bytes(H, Hkv, L, d, w, bound))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        mc, ks = c_int(0), c_int64(0)
        offs = (c_size_t * 3)()
        _check("kvc_adakv_prefill", lib().kvc_adakv_prefill(
            q.data_ptr(), q.stride(1), q.stride(2), k.data_ptr(), k.stride(1), k.stride(2),
            v.data_ptr(), v.stride(1), v.stride(2), H, Hkv, L, d, w, dtype_code(k.dtype), kernel_size,
            pool_code(pooling), base_capacity, one_minus_floor, int(base_capacity * floor_ratio),
            int(bool(normalize)), impl, flat_k.data_ptr(), flat_v.data_ptr(), out_rows, cap.data_ptr(),
            head_lens.data_ptr(), cu_klen.data_ptr(), host.data_ptr(), bound, ws.data_ptr(), nbytes,
            ctypes.byref(mc), ctypes.byref(ks), offs, _stream(dev)))
    max_cap, klen_sum = mc.value, ks.value
    debug = None
    if want_debug:
        debug = dict(scores=ws[offs[0]: offs[0] + H * n * e].view(k.dtype).view(1, H, n),
                     pooled=ws[offs[1]: offs[1] + H * n * e].view(k.dtype).view(1, H, n), cap=cap,
                     indices=ws[offs[2]: offs[2] + H * max_cap * 4].view(torch.int32).view(H, max_cap))
    return flat_k[:klen_sum], flat_v[:klen_sum], cap, head_lens, cu_klen, max_cap, klen_sum, debug


def adakv_key_histogram(pooled: torch.Tensor, sums: torch.Tensor, normalize: bool) -> torch.Tensor:
    """int32 [H,65536] histogram of each head's ratio-normalised scores over the ordered 16-bit key."""
    dev = _require_cuda(pooled, sums)
    H, n = pooled.shape
    hist = torch.empty((H, 65536), dtype=torch.int32, device=dev)
    with _on_device(dev):
        _check("kvc_adakv_key_histogram", lib().kvc_adakv_key_histogram(
            pooled.contiguous().data_ptr(), sums.data_ptr(), H, n, dtype_code(pooled.dtype), int(bool(normalize)),
            hist.data_ptr(), _stream(dev)))
    return hist


def adakv_hist_coarse(pooled: torch.Tensor, sums: torch.Tensor, normalize: bool):
    """(akeys uint16 [H,n] as int16 storage, hist1 int32 [H,4096]): first counting phase of the AdaKV allocation."""
    dev = _require_cuda(pooled, sums)
    H, n = pooled.shape
    akeys = torch.empty((H, n), dtype=torch.int16, device=dev)
    hist1 = torch.empty((H, 4096), dtype=torch.int32, device=dev)
    with _on_device(dev):
        _check("kvc_adakv_hist_coarse", lib().kvc_adakv_hist_coarse(
            pooled.contiguous().data_ptr(), sums.data_ptr(), H, n, dtype_code(pooled.dtype), int(bool(normalize)),
            akeys.data_ptr(), hist1.data_ptr(), _stream(dev)))
    return akeys, hist1


def adakv_hist_fine(akeys: torch.Tensor, hist1: torch.Tensor, thr_bin: torch.Tensor):
    """(hist2 int32 [H,16], above int32 [H]) for the DEVICE scalar ``thr_bin`` (int32): second counting phase."""
    dev = _require_cuda(akeys, hist1, thr_bin)
    if thr_bin.dtype != torch.int32 or thr_bin.numel() != 1:
        raise TypeError("thr_bin must be one int32 on the device")
    H, n = akeys.shape
    hist2 = torch.empty((H, 16), dtype=torch.int32, device=dev)
    above = torch.empty(H, dtype=torch.int32, device=dev)
    with _on_device(dev):
        _check("kvc_adakv_hist_fine", lib().kvc_adakv_hist_fine(
            akeys.data_ptr(), hist1.data_ptr(), H, n, thr_bin.data_ptr(), hist2.data_ptr(), above.data_ptr(),
            _stream(dev)))
    return hist2, above


def think_prune(k_compressed: torch.Tensor, query_states: torch.Tensor, recent_size: int, ratio: float):
    """ThinK channel pruning of the compressed keys (``key_pruner_query_driven``, pyramidkv_utils.py:13-26):
    (kv_pruned [1, B*H, rows-recent, d-k], kv_recent [B,H,recent,d] (a view), mask bool [B,H,d], True = kept)."""
    dev = _require_cuda(k_compressed, query_states)
    B, H, rows, d = k_compressed.shape
    if query_states.shape[0] != B or query_states.shape[1] != H or query_states.dtype != k_compressed.dtype:
        raise ValueError("query_states must be [B,H,L,d] with the heads and dtype of the compressed keys")
    L = query_states.shape[2]
    kc = k_compressed.contiguous()
    q = _rowmajor_last(query_states, 1)
    n_drop = int(d * ratio)                                             # :15
    recent = min(recent_size, rows)
    pruned = torch.empty((1, B * H, rows - recent, d - n_drop), dtype=kc.dtype, device=dev)
    mask = torch.empty((B, H, d), dtype=torch.bool, device=dev)
    with _on_device(dev):
        nbytes = max(256, lib().kvc_think_prune_workspace_bytes(B, H, rows, d))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _check("kvc_think_prune", lib().kvc_think_prune(
            kc.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), B, H, L, rows, d, recent, n_drop,
            dtype_code(kc.dtype), pruned.data_ptr(), mask.data_ptr(), ws.data_ptr(), nbytes, _stream(dev)))
    return pruned, kc[:, :, rows - recent:, :], mask


def update_flatten_view(cache: torch.Tensor, state: torch.Tensor, headlens: torch.Tensor,
                        cu_headlens: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``tiny_api_cuda.update_flatten_view`` (csrc/csrc/cuda_api.cu:55-85)."""
    if headlens.dtype != torch.int32:
        raise RuntimeError("expected headlens to be int32")          # cuda_api.cu:56
    if cu_headlens.dtype != torch.int32:
        raise RuntimeError("expected cu_dst_pos to be int32")        # cuda_api.cu:57
    dev = _require_cuda(cache, state, headlens, cu_headlens)
    dtype_code(cache.dtype)
    total, d = cache.shape
    H = headlens.shape[0]
    if d > 128:
        raise RuntimeError("num threads should larger than head dim")  # cuda_api.cu:74
    cache = cache.contiguous()
    state = state.contiguous()
    out = torch.empty((total + H, d), dtype=cache.dtype, device=dev)
    with _on_device(dev):
        _check("kvc_update_flatten_view", lib().kvc_update_flatten_view(
            out.data_ptr(), cache.data_ptr(), state.data_ptr(), headlens.data_ptr(), cu_headlens.data_ptr(),
            H, d, cache.element_size(), total, _stream(dev)))
    return out


def append_rows(k_buf: torch.Tensor, v_buf: torch.Tensor, new_k: torch.Tensor, new_v: torch.Tensor,
                cu_cap: torch.Tensor, head_lens: torch.Tensor) -> None:
    """In place: new_k[h] / new_v[h] -> row cu_cap[h] + head_lens[h] of the slack-layout buffers (O(H*d))."""
    dev = _require_cuda(k_buf, v_buf, new_k, new_v, cu_cap, head_lens)
    if cu_cap.dtype != torch.int32 or head_lens.dtype != torch.int32:
        raise RuntimeError("expected cu_cap / head_lens to be int32")
    dtype_code(k_buf.dtype)
    H, d = new_k.shape
    if not (k_buf.is_contiguous() and v_buf.is_contiguous()):
        raise RuntimeError("slack-layout cache buffers must be contiguous")
    with _on_device(dev):
        _check("kvc_append_rows", lib().kvc_append_rows(
            k_buf.data_ptr(), v_buf.data_ptr(), new_k.contiguous().data_ptr(), new_v.contiguous().data_ptr(),
            cu_cap.data_ptr(), head_lens.data_ptr(), H, d, k_buf.element_size(), _stream(dev)))


def flatten_regrow(src: torch.Tensor, head_lens: torch.Tensor, cu_src: torch.Tensor, cu_dst: torch.Tensor,
                   total_dst_rows: int) -> torch.Tensor:
    """New [total_dst_rows, d] buffer with head h's head_lens[h] rows moved from src row cu_src[h] to row cu_dst[h]."""
    dev = _require_cuda(src, head_lens, cu_src, cu_dst)
    dtype_code(src.dtype)
    H, d = head_lens.shape[0], src.shape[1]
    dst = torch.empty((total_dst_rows, d), dtype=src.dtype, device=dev)
    with _on_device(dev):
        _check("kvc_flatten_regrow", lib().kvc_flatten_regrow(
            dst.data_ptr(), src.contiguous().data_ptr(), head_lens.data_ptr(), cu_src.data_ptr(), cu_dst.data_ptr(),
            H, d, src.element_size(), src.shape[0], _stream(dev)))
    return dst


def launch_count() -> int:
    """Kernels launched by libkvc_b200 in this process so far."""
    return int(lib().kvc_launch_count())


def sm_count() -> int:
    n = c_int(0)
    _check("kvc_sm_count", lib().kvc_sm_count(ctypes.byref(n)))
    return n.value
