"""Host-buffer entry to the compression path: pinned host K/V/Q-window in, compressed K'/V' out.

This is the call a caller without device-resident tensors makes (offloaded prefill, a remote
prefill worker handing KV over host memory, the ``e2e`` leg of ``bench.py``).  Per layer it copies
K and the last ``window_size`` query rows to the device (the only query rows the window-scored
policies read, reference ``pyramidkv_utils.py:317``), runs the cluster's CUDA path, and copies the
compressed rows back.  V never moves as a whole: the path reads V only at the selected rows, so the
gather kernel loads those rows directly from the pinned host buffer (``zero_copy_v``; 4x fewer V
bytes over PCIe at cfg-2).  Copies run on their own streams with double-buffered staging so that
layer l+1's host->device transfer and layer l-1's device->host transfer overlap layer l's kernels.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class HostCompressor:
    def __init__(self, clusters: Sequence, num_key_value_groups: int, device: torch.device, zero_copy_v: bool = True):
        self.zero_copy_v = zero_copy_v
        self.clusters = list(clusters)
        self.groups = num_key_value_groups
        self.device = device
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self._stage = None       # two (K, V, Qwin) device staging sets
        self._out = None         # pinned host outputs per layer
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staging(self, k: torch.Tensor, q: torch.Tensor):
        if self._stage is None or self._stage[0][0].shape != k.shape or self._stage[0][2].shape != q.shape:
            self._stage = [(torch.empty(k.shape, dtype=k.dtype, device=self.device),
                            torch.empty(k.shape, dtype=k.dtype, device=self.device),
                            torch.empty(q.shape, dtype=q.dtype, device=self.device)) for _ in range(2)]
        return self._stage

    def run(self, host_k: List[torch.Tensor], host_q_window: List[torch.Tensor], host_v: List[torch.Tensor],
            seq_len: int):
        """host_k/host_v: per-layer pinned ``[B,Hkv,L,d]``; host_q_window: per-layer pinned ``[B,Hq,w,d]``.
        Returns per-layer (K', V') pinned host tensors (reused across calls)."""
        n = len(self.clusters)
        cur = torch.cuda.current_stream(self.device)
        stage = self._staging(host_k[0], host_q_window[0])
        if self._out is None:
            self._out = [None] * n
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_done = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        keep = []
        h2d = d2h = 0
        self.s_in.wait_stream(cur)
        for l in range(n):
            dk, dv, dq = stage[l % 2]
            with torch.cuda.stream(self.s_in):
                if l >= 2:
                    self.s_in.wait_event(ev_done[l - 2])           # staging set is free again ...
                    self.s_in.wait_event(ev_out[l - 2])            # ... also for the output copies: on the pass-through
                                                                   # branch K' IS the staging buffer
                dk.copy_(host_k[l], non_blocking=True)
                if not self.zero_copy_v:
                    dv.copy_(host_v[l], non_blocking=True)
                dq.copy_(host_q_window[l], non_blocking=True)
                ev_in[l].record(self.s_in)
            cur.wait_event(ev_in[l])
            # select + gather stay on `cur` here: the staging buffers and the output copies are ordered by
            # events on `cur`, and the step is PCIe-bound anyway (the tail is ~2 % of a layer's copy time)
            tail_mode, self.clusters[l].async_tail = getattr(self.clusters[l], "async_tail", False), False
            try:
                kc, vc = self.clusters[l].update_kv_window(dk, dq, host_v[l] if self.zero_copy_v else dv, self.groups,
                                                           seq_len)
            finally:
                self.clusters[l].async_tail = tail_mode

            esz = host_k[l].element_size()
            # bytes that cross PCIe towards the device: all of K, the query window, and either all of V or
            # only the V rows the gather kernel reads
            h2d += host_k[l].numel() * esz + host_q_window[l].numel() * esz
            h2d += vc.numel() * esz if self.zero_copy_v and vc is not host_v[l] else host_v[l].numel() * esz
            ev_done[l].record(cur)
            if self._out[l] is None or self._out[l][0].shape != kc.shape:
                self._out[l] = (torch.empty(kc.shape, dtype=kc.dtype, pin_memory=True),
                                torch.empty(vc.shape, dtype=vc.dtype, pin_memory=True))
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done[l])
                self._out[l][0].copy_(kc, non_blocking=True)
                self._out[l][1].copy_(vc, non_blocking=True)
                ev_out[l].record(self.s_out)
            keep.append((kc, vc))
            d2h += 2 * kc.numel() * kc.element_size()
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        del keep
        return self._out
