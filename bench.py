#!/usr/bin/env python
"""bench.py — KV tokens compressed / second on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (default ``pyramidkv_llama3_8b_32k`` = BASELINE.json configs[1]): PyramidKV compression of a
32-layer Llama-3-8B prefill (32 query / 8 KV heads, d=128, bf16) of 32 768 tokens to budget 2048
(window 32, avg-pool 7).  One "step" = ``update_kv`` of all 32 layers on synthetic N(0,1) K/V/Q.
Metric = B*L / time-per-step, whole job over all ranks.

Legs printed in ONE JSON line (rank 0):
  value        device-resident inputs, CUDA events around exactly K steps, max over ranks
  sync_ms_per_step  the same step with the drop-in DEFAULTS (everything on the caller's stream)
  e2e          same metric through the host-buffer API: pinned host K/V/Q-window -> device,
               compress, compressed K'/V' -> host, all inside the timed region
  roofline     dominant kernel (time share measured with CUDA events inside the timed region):
               algorithmic bytes per launch / its mean duration vs MEASURED_PEAKS.json
  cpu_baseline the reference's own update_kv (oracle/_ref copy of the unmodified file when present, else the
               oracle port) on the host cores, bounded sample, rank 0, N=1
  cfg5         BASELINE configs[4] as north_star states it: ONE Llama-3-70B prefill of 131 072 tokens, its 8 KV heads
               sharded 8/N per rank, 80 layers, HBM-cold inputs: strong scaling, no data-path collective
  cfg4         (N=1) BASELINE configs[3]: AdaKV at 64K, one layer's update_kv and the decode-side append
  cfg1, cfg3   (N=1) BASELINE configs[0] (SnapKV, one layer at 4K) and configs[2] (H2O / StreamingLLM, Mistral-7B shapes,
               32K, batch 8, one layer each)
  checks       (N>1) ShardedAdaKV's budget exchange over NCCL gives the single-GPU capacities
``--impl reference`` times the reference's CPU path alone on the same workload (bounded sample per step).

Multi-GPU (torchrun, one rank per GPU): `value` shards the (batch, KV-head) units of an N-request batch over the
ranks, 8 units per rank as at N=1, with no data-path collective ("scaling": "weak"); `cfg5` is the strong-scaling
companion.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "pyramidkv_llama3_8b_32k": dict(method="pyramidkv", B=1, Hq=32, Hkv=8, d=128, L=32768, layers=32, window=32,
                                    cap=2048, kernel=7, pooling="avgpool", dtype="bf16"),
    # configs[0]: the reference's own CPU-runnable case (one layer)
    "snapkv_llama3_8b_4k_1layer": dict(method="snapkv", B=1, Hq=32, Hkv=8, d=128, L=4096, layers=1, window=32,
                                       cap=128, kernel=7, pooling="avgpool", dtype="bf16"),
    # configs[4] per-GPU shard at 8 GPUs is 1 KV head; at N=1 this is the whole 70B layer stack
    "pyramidkv_llama3_70b_128k": dict(method="pyramidkv", B=1, Hq=64, Hkv=8, d=128, L=131072, layers=80, window=32,
                                      cap=2048, kernel=7, pooling="avgpool", dtype="bf16"),
}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes
# ------------------------------------------------------------------------------------------------
def kept_rows(wl, layer):
    from kvcache_factory_b200 import pyramid_budget
    if wl["method"] == "pyramidkv":
        base = wl["cap"] - wl["window"]
        if wl["L"] < wl["cap"]:
            return wl["L"]
        if wl["L"] < 2 * base:
            return wl["cap"]
        return pyramid_budget(wl["L"], wl["window"], wl["cap"], wl["layers"], layer) + wl["window"]
    return wl["cap"]


def alg_bytes_layer(wl, layer, B=None, Hq=None, Hkv=None):
    e = 2
    B = B or wl["B"]; Hq = Hq or wl["Hq"]; Hkv = Hkv or wl["Hkv"]
    score = B * Hkv * wl["L"] * wl["d"] * e + B * Hq * wl["window"] * wl["d"] * e
    C = kept_rows(wl, layer)
    gather = 4 * B * Hq * C * wl["d"] * e
    if wl["method"] == "streamingllm":
        score = 0
    return score, gather


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def load_reference_module():
    """The reference's own policy file, byte-copied into oracle/_ref/ by oracle/build_ref.py (git-ignored; it
    travels to the GPU box with the snapshot).  None when it is not there: the oracle port is timed instead."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_pyramidkv_utils.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    try:
        spec = importlib.util.spec_from_file_location("ref_pyramidkv_utils", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def cpu_reference(wl, sample_layers, repeats=1):
    """Time the reference's update_kv on the host cores on `sample_layers` -- the UNMODIFIED reference file when
    oracle/_ref holds it ("reference"), else the CPU restatement oracle/kvc_oracle.py ("port") -- with repeat_kv
    first, as at every reference call site (llama_model.py:379-380).  Returns (tokens/s extrapolated to all layers,
    seconds per layer, cores, kind)."""
    import contextlib
    import io
    import torch
    from kvcache_factory_b200.synth import Shape, make_qkv
    from oracle import kvc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = load_reference_module()
    shape = Shape(wl["B"], wl["Hq"], wl["Hkv"], wl["L"], wl["d"])
    dtype = torch.bfloat16 if wl["dtype"] == "bf16" else torch.float16
    groups = wl["Hq"] // wl["Hkv"]
    times = []
    for layer in sample_layers:
        Q, K, V = make_qkv(shape, dtype, layer_idx=layer, window=wl["window"], q_window_only=True)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            if ref is not None:
                with contextlib.redirect_stdout(io.StringIO()):        # the reference prints on every call (:217,312)
                    Kr, Vr = ref.repeat_kv(K, groups), ref.repeat_kv(V, groups)
                    if wl["method"] == "pyramidkv":
                        cl = ref.PyramidKVCluster(num_hidden_layers=wl["layers"], window_size=wl["window"],
                                                  max_capacity_prompt=wl["cap"], kernel_size=wl["kernel"],
                                                  pooling=wl["pooling"], layer_idx=layer)
                    else:
                        cl = ref.SnapKVCluster(window_size=wl["window"], max_capacity_prompt=wl["cap"],
                                               kernel_size=wl["kernel"], pooling=wl["pooling"])
                    cl.update_kv(Kr, Q, Vr, None, groups)
            elif wl["method"] == "pyramidkv":
                O.pyramidkv_update_kv(K, Q, V, layer, wl["layers"], wl["window"], wl["cap"], wl["kernel"], wl["pooling"])
            else:
                O.snapkv_update_kv(K, Q, V, wl["window"], wl["cap"], wl["kernel"], wl["pooling"])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        times.append(best)
    per_layer = sum(times) / len(times)
    return wl["B"] * wl["L"] / (per_layer * wl["layers"]), per_layer, cores, ("reference" if ref is not None else "port")


def run_reference(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    n = wl["layers"]
    sample = sorted(set([0, n // 2, n - 1]))[: max(1, min(3, n))]
    for _ in range(min(args.warmup, 1)):
        cpu_reference(wl, sample[:1])
    vals = []
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        v, per_layer, cores, kind = cpu_reference(wl, sample)
        vals.append(v)
    ms_step = (time.perf_counter() - t_all) * 1e3 / max(1, args.steps)
    value = statistics.median(vals)
    what = ("the unmodified reference pyramidkv_utils.py (oracle/_ref copy), repeat_kv + update_kv on torch CPU"
            if kind == "reference" else "oracle/kvc_oracle.py (torch CPU restatement of update_kv incl. repeat_kv)")
    desc = f"{what}, layers {sample} of {n} at full L={wl['L']}, mean per-layer time x {n} layers"
    line = {
        "impl": "reference", "metric": "kv_tokens_compressed_per_s", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": config_of(args.workload, wl, args.gpus),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_of(name, wl, world):
    return {"workload": name, "method": wl["method"], "heads": f"{wl['Hq']}q/{wl['Hkv']}kv", "head_dim": wl["d"],
            "seq_len": wl["L"], "layers": wl["layers"], "window_size": wl["window"],
            "max_capacity_prompt": wl["cap"], "kernel_size": wl["kernel"], "pooling": wl["pooling"],
            "batch_per_gpu": wl["B"], "sharding": f"{wl['B'] * wl['Hkv']} (batch,kv-head) units per GPU x {world}",
            "l2": "inputs larger than L2 (4 GiB of K/V per step per GPU)"}



# ------------------------------------------------------------------------------------------------
# side blocks of the JSON line: cfg5 (strong scaling over KV heads), cfg4 (AdaKV), NCCL check
# ------------------------------------------------------------------------------------------------
def _graph_of(torch, fn):
    """Capture fn() (which may fan out over the library's side streams) into one CUDA graph on a high-priority stream."""
    side = torch.cuda.Stream(priority=-1)
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        with torch.cuda.graph(g, stream=side):
            keep = fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    return g, keep


def _time_replays(torch, g, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def measure_cfg5(torch, dist, dev, rank, world, steps, async_tail, async_scores, peak_gbs):
    """BASELINE configs[4] / SURVEY.md section 8(e): ONE prefill of Llama-3-70B shapes (64 q / 8 KV heads, 80 layers)
    at L = 131072, PyramidKV to cap 2048; rank r owns KV heads [r*8/N, (r+1)*8/N) and their query groups (as under
    head-parallel tensor parallelism), no data-path collective.  Strong scaling: the job is the same at every N.
    K/V of the 80 layers rotate over a pool of 6 layer buffers (> L2), so every layer reads HBM-cold keys."""
    from kvcache_factory_b200 import PyramidKVCluster, wait_for_compression
    from kvcache_factory_b200.sharding import kv_head_slice
    wl = WORKLOADS["pyramidkv_llama3_70b_128k"]
    if wl["Hkv"] % world:
        return {"skipped": f"8 KV heads do not split over {world} ranks"}
    L, d, w, nl, g = wl["L"], wl["d"], wl["window"], wl["layers"], wl["Hq"] // wl["Hkv"]

    def run(heads, nsteps):
        gen = torch.Generator(device=dev)
        gen.manual_seed(777 + rank)
        pool = 6
        Ks = [torch.randn(1, heads, L, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(pool)]
        Vs = [torch.randn(1, heads, L, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(pool)]
        Qw = [torch.randn(1, w, heads * g, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16).transpose(1, 2)
              for _ in range(nl)]
        cls = []
        for l in range(nl):
            c = PyramidKVCluster(num_hidden_layers=nl, window_size=w, max_capacity_prompt=wl["cap"], kernel_size=wl["kernel"],
                                 pooling=wl["pooling"], layer_idx=l)
            c.async_tail, c.async_scores = async_tail, async_scores
            cls.append(c)

        def step():
            out = None
            for l in range(nl):
                out = cls[l].update_kv_window(Ks[l % pool], Qw[l], Vs[l % pool], g, L)
            wait_for_compression(dev)
            return out

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        gr, keep = _graph_of(torch, step)
        gr.replay()
        torch.cuda.synchronize()
        ms = _time_replays(torch, gr, nsteps)
        del gr, keep, Ks, Vs, Qw
        torch.cuda.empty_cache()
        return ms

    heads = wl["Hkv"] // world
    sl = kv_head_slice(wl["Hkv"], world, rank)
    if world > 1:
        dist.barrier()
    ms = run(heads, max(3, min(steps, 10)))
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    by = sum(sum(alg_bytes_layer(wl, l, Hq=heads * g, Hkv=heads)) for l in range(nl))
    out = {"workload": "pyramidkv_llama3_70b_128k", "scaling": "strong", "kv_heads_per_rank": heads,
           "kv_heads_of_rank0": [sl.start, sl.stop], "layers": nl, "seq_len": L, "ms_per_step": ms,
           "tokens_per_s": L / (ms * 1e-3), "alg_bytes_per_rank": by, "hbm_frac_per_rank": by / (ms * 1e-3) / 1e9 / peak_gbs,
           "inputs": "K/V rotate over 6 layer buffers per rank (> L2): HBM-cold", "collectives": 0}
    if world > 1:
        # the same job on ONE GPU (all 8 heads), measured here by rank 0 so that the efficiency has its denominator
        n1 = run(wl["Hkv"], 3) if rank == 0 else 0.0
        dist.barrier()
        out["n1_ms_per_step_same_box"] = n1
        out["strong_scaling_efficiency"] = (n1 / (world * ms)) if rank == 0 and ms > 0 else None
    else:
        out["strong_scaling_efficiency"] = 1.0
    return out


def measure_cfg4(torch, dev, peak_gbs):
    """BASELINE configs[3]: AdaKV (cap 2048, floor 0.2, maxpool 7, normalize) on Llama-3-8B shapes at L = 65536: one
    layer's update_kv (eager calls, inputs rotate over 3 buffer sets > L2) and the decode-side flattened append."""
    from kvcache_factory_b200 import AdaKVCluster, _native
    L, Hq, Hkv, d, w, cap = 65536, 32, 8, 128, 32, 2048
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    Ks = [torch.randn(1, Hkv, L, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(3)]
    Vs = [torch.randn(1, Hkv, L, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(3)]
    Q = torch.zeros(1, L, Hq, d, device=dev, dtype=torch.bfloat16)
    Q[:, L - w:] = torch.randn(1, w, Hq, d, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    Q = Q.transpose(1, 2)
    cl = AdaKVCluster(window_size=w, kernel_size=7, pooling="maxpool", max_capacity_prompt=cap, floor=0.2, normalize=True)
    for i in range(3):
        fk, fv = cl.update_kv(Ks[i % 3], Q, Vs[i % 3])
    torch.cuda.synchronize()
    ts = []
    for i in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fk, fv = cl.update_kv(Ks[i % 3], Q, Vs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    us = statistics.median(ts) * 1e3
    rows = int(fk.shape[0])
    by = Hkv * L * d * 2 + Hq * w * d * 2 + 4 * rows * d * 2
    # decode append: K and V of one layer-step, replayed from a graph over 3 rotating caches
    caches = [torch.randn(rows, d, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(6)]
    state = torch.randn(Hq, d, device=dev, dtype=torch.float32).to(torch.bfloat16)

    def app():
        return [_native.update_flatten_view(c, state, cl.head_lens, cl.cu_klen) for c in caches]

    g, keep = _graph_of(torch, app)
    app_us = _time_replays(torch, g, 10) * 1e3 / len(caches)
    app_by = (2 * rows + Hq) * d * 2
    out = {"workload": "adakv_llama3_8b_64k", "prefill_us_per_layer": us, "flat_rows": rows, "alg_bytes": by,
           "prefill_hbm_frac": by / (us * 1e-6) / 1e9 / peak_gbs,
           "decode_append_us_per_tensor": app_us, "decode_append_hbm_frac": app_by / (app_us * 1e-6) / 1e9 / peak_gbs,
           "note": "prefill: eager update_kv incl. the reference's host read of the capacities; append: CUDA-graph replay"}
    del g, keep, caches, Ks, Vs, Q
    torch.cuda.empty_cache()
    return out


def _event_us(torch, fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def measure_cfg1(torch, dev, peak_gbs):
    """BASELINE configs[0]: SnapKV update_kv on one Llama-3-8B layer at L = 4096 (cap 128, avgpool 7): the drop-in call,
    eager (one library call, three launches) and replayed from a CUDA graph over 6 rotating K/V sets."""
    from kvcache_factory_b200 import SnapKVCluster
    B, Hq, Hkv, L, d, w, cap = 1, 32, 8, 4096, 128, 32, 128
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    r = lambda *sh: torch.randn(*sh, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    Ks, Vs = [r(B, Hkv, L, d) for _ in range(6)], [r(B, Hkv, L, d) for _ in range(6)]
    Q = r(B, Hq, L, d)
    cl = SnapKVCluster(w, cap, 7, "avgpool")
    turn = [0]

    def call():
        i = turn[0] = (turn[0] + 1) % 6
        return cl.update_kv(Ks[i], Q, Vs[i], None, Hq // Hkv)

    eager = _event_us(torch, call, 30, 5)
    g, keep = _graph_of(torch, lambda: [call() for _ in range(12)])
    rep = _time_replays(torch, g, 10) * 1e3 / 12
    del keep
    alg = B * Hkv * L * d * 2 + B * Hq * w * d * 2 + 4 * B * Hq * cap * d * 2
    return {"workload": "snapkv_llama3_8b_4k_one_layer", "eager_us_per_layer": eager, "graph_us_per_layer": rep,
            "alg_bytes": alg, "graph_hbm_frac": alg / (rep * 1e-6) / 1e9 / peak_gbs,
            "note": "12.8 MB per layer: the three launches are latency-, not bandwidth-bound; inputs rotate over 6 sets "
                    "(101 MB) but still fit the 126 MB L2"}


def measure_cfg3(torch, dev, peak_gbs, peak_tf):
    """BASELINE configs[2]: H2O and StreamingLLM on Mistral-7B shapes (32 q / 8 kv heads), L = 32768, budget 1024,
    batch 8: one layer's update_kv each."""
    from kvcache_factory_b200 import H2OKVCluster, StreamingLLMKVCluster
    B, Hq, Hkv, L, d, w, cap = 8, 32, 8, 32768, 128, 32, 1024
    gen = torch.Generator(device=dev)
    gen.manual_seed(12)
    r = lambda *sh: torch.randn(*sh, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    K, V, Q = r(B, Hkv, L, d), r(B, Hkv, L, d), r(B, Hq, L, d)
    st = StreamingLLMKVCluster(w, cap, 7, "avgpool")
    s_us = _event_us(torch, lambda: st.update_kv(K, Q, V, None, Hq // Hkv), 10, 3)
    s_alg = 4 * B * Hq * cap * d * 2
    h2o = H2OKVCluster(w, cap, 7, "avgpool")
    h_us = _event_us(torch, lambda: h2o.update_kv(K, Q, V, None, Hq // Hkv), 2, 1)
    flops = 4.0 * B * Hq * L * L * d          # two passes of Q.K^T over all L x L logits per head
    out = {"workload": "h2o_streamingllm_mistral7b_32k_b8_one_layer",
           "streamingllm_us_per_layer": s_us, "streamingllm_hbm_frac": s_alg / (s_us * 1e-6) / 1e9 / peak_gbs,
           "h2o_ms_per_layer": h_us * 1e-3, "h2o_tflops": flops / (h_us * 1e-6) / 1e12,
           "h2o_exps_per_s": 2.0 * B * Hq * L * L / (h_us * 1e-6),
           "note": "H2O scores all L x L logits in two tcgen05 passes without materialising them: bound by the SFU "
                   "(one exp per logit and pass, 16 per clock and SM = 4.65e12/s at 1965 MHz), not by the tensor pipe"}
    if peak_tf:
        out["h2o_tensor_frac"] = out["h2o_tflops"] / peak_tf
    out["h2o_sfu_frac"] = out["h2o_exps_per_s"] / (16 * 148 * 1.965e9)
    return out


def check_sharded_adakv(torch, dist, dev, rank, world):
    """AdaKV's one exchange step on real NCCL: every rank runs ShardedAdaKV on its KV-head shard of the same seeded
    input; the capacities must equal those of the single-GPU AdaKVCluster over all heads (computed by every rank)."""
    from kvcache_factory_b200 import AdaKVCluster
    from kvcache_factory_b200.sharding import ShardedAdaKV, shard_qkv
    from kvcache_factory_b200.synth import Shape, make_qkv
    Hkv = 8
    if Hkv % world:
        return {"skipped": f"8 KV heads do not split over {world} ranks"}
    shape = Shape(1, 32, Hkv, 8192, 128)
    Q, K, V = make_qkv(shape, torch.bfloat16, layer_idx=11, dist="peaked", window=32, q_window_only=True)
    Qd, Kd, Vd = Q.to(dev), K.to(dev), V.to(dev)
    mk = lambda: AdaKVCluster(window_size=32, kernel_size=7, pooling="maxpool", max_capacity_prompt=1024, floor=0.2,
                              normalize=True)
    full = mk()
    fk, fv = full.update_kv(Kd, Qd, Vd)
    sh = ShardedAdaKV(mk())
    q, k, v = shard_qkv(Qd, Kd, Vd, world, rank)
    sk, sv = sh.update_kv(k.contiguous(), q.contiguous(), v.contiguous())
    torch.cuda.synchronize()
    per = 32 // world
    want = full.head_lens[rank * per:(rank + 1) * per]
    ok_caps = bool(torch.equal(sh.cluster.head_lens, want)) and bool(torch.equal(sh.cap_all.to(dev) + 32, full.head_lens))
    cu = full.cu_klen.tolist()
    ok_rows = bool(torch.equal(sk, fk[cu[rank * per]:cu[(rank + 1) * per]]))
    t = torch.tensor([int(ok_caps), int(ok_rows)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"sharded_adakv_capacities_equal_single_gpu": bool(t[0].item()), "sharded_adakv_rows_equal_single_gpu": bool(t[1].item()),
            "backend": dist.get_backend(), "world": world}

# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # One process per GPU and one stream of scoring kernels: the fused scoring kernel may be launched
    # without the cooperative attribute so that its CTAs start while the previous layer's asynchronous
    # select/gather tail is still draining (see scores_fused.cu).  Must be set before the library loads.
    os.environ.setdefault("KVC_FUSED_COOP", "0")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from kvcache_factory_b200 import PyramidKVCluster, SnapKVCluster, _native
    from kvcache_factory_b200.host_api import HostCompressor

    B, Hq, Hkv, L, d, w, nl = wl["B"], wl["Hq"], wl["Hkv"], wl["L"], wl["d"], wl["window"], wl["layers"]
    dtype = torch.bfloat16 if wl["dtype"] == "bf16" else torch.float16
    impl = {"auto": _native.IMPL_AUTO, "mma": _native.IMPL_MMA_SYNC, "tc": _native.IMPL_TCGEN05,
            "fused": _native.IMPL_TCGEN05_FUSED, "half": _native.IMPL_TCGEN05_HALF, "dyn": _native.IMPL_TCGEN05_DYN}[args.score_impl]

    # synthetic inputs, device resident; each rank draws its own shard of the N-request batch
    g = torch.Generator(device=dev)
    Ks, Vs, Qs, clusters = [], [], [], []
    q_window_only = nl * B * L * Hq * d * 2 > (32 << 30)
    need = 2 * nl * B * Hkv * L * d * 2 + (0 if q_window_only else nl * B * L * Hq * d * 2)
    if need > 0.8 * torch.cuda.get_device_properties(dev).total_memory:
        raise SystemExit(f"workload {args.workload} needs {need / 2**30:.0f} GiB of device-resident inputs per GPU: shard it over more GPUs")
    for layer in range(nl):
        g.manual_seed(1234 + layer + 1000 * rank)
        Ks.append(torch.randn(B, Hkv, L, d, generator=g, device=dev, dtype=torch.float32).to(dtype))
        Vs.append(torch.randn(B, Hkv, L, d, generator=g, device=dev, dtype=torch.float32).to(dtype))
        # query_states as HF hands it over: a [B,L,Hq,d] buffer viewed as [B,Hq,L,d]; only the window is read
        if q_window_only:   # 70B at 128K: 80 full-length query tensors would be 160 GiB; the path reads the window only
            Qs.append(torch.randn(B, w, Hq, d, generator=g, device=dev, dtype=torch.float32).to(dtype).transpose(1, 2))
        else:
            qbuf = torch.zeros(B, L, Hq, d, device=dev, dtype=dtype)
            qbuf[:, L - w:] = torch.randn(B, w, Hq, d, generator=g, device=dev, dtype=torch.float32).to(dtype)
            Qs.append(qbuf.transpose(1, 2))
        if wl["method"] == "pyramidkv":
            cl = PyramidKVCluster(num_hidden_layers=nl, window_size=w, max_capacity_prompt=wl["cap"],
                                  kernel_size=wl["kernel"], pooling=wl["pooling"], layer_idx=layer)
        else:
            cl = SnapKVCluster(w, wl["cap"], wl["kernel"], wl["pooling"])
        cl.score_impl = impl
        cl.async_tail = not args.sync_tail
        cl.async_scores = not args.sync_tail and not args.sync_scores
        clusters.append(cl)
    groups = Hq // Hkv

    from kvcache_factory_b200 import wait_for_compression

    def step():
        out = None
        for layer in range(nl):
            if q_window_only:
                out = clusters[layer].update_kv_window(Ks[layer], Qs[layer], Vs[layer], groups, L)
            else:
                out = clusters[layer].update_kv(Ks[layer], Qs[layer], Vs[layer], None, groups)
        wait_for_compression(dev)  # join the asynchronous select/gather tails: all work ends inside the step
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, optionally replayed as one CUDA graph per step ------------------
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        # the scoring kernels run on a HIGH-priority stream, the select/gather tails on the library's
        # default-priority side stream: the block scheduler then places the next layer's scoring CTAs
        # first and lets the tail fill what is left of each SM (kernel-node priorities are captured)
        side = torch.cuda.Stream(priority=-1)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
            with torch.cuda.graph(graph, stream=side):
                graph_out = step()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    run_step = graph.replay if graph is not None else step
    used_graph = graph is not None

    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _native.launch_count() - launches0
    if graph is not None:  # graph replays do not pass through the library: count what was captured
        l0 = _native.launch_count(); step(); launches = (_native.launch_count() - l0) * args.steps
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * L / (ms_per_step * 1e-3)

    # ---- the same step with the drop-in defaults: every kernel on the caller's stream, AUTO scoring kernel ----
    sync_ms = None
    if not args.sync_tail:
        saved = [(c.async_tail, c.async_scores) for c in clusters]
        for c in clusters:
            c.async_tail = c.async_scores = False
        try:
            gs, _keep = _graph_of(torch, step)
            barrier()
            sync_ms = _time_replays(torch, gs, max(3, min(args.steps, 10)))
            del gs, _keep
        finally:
            for c, (a, b_) in zip(clusters, saved):
                c.async_tail, c.async_scores = a, b_

    # ---- per-kernel time, live: each phase's 32 launches (one per layer) replayed as its own CUDA
    # graph and bracketed by CUDA events on the launching stream (no host gaps inside the bracket) ----
    def phase_graph(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn()
            with torch.cuda.graph(g_, stream=side):
                keep = fn()
        torch.cuda.current_stream().wait_stream(side)
        return g_, keep

    def time_graph(g_, reps):
        g_.replay()
        barrier()  # ranks time the same phase at the same time (no neighbour's host copies in the bracket)
        ts = []
        for _ in range(reps):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g_.replay()
            b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        return statistics.median(ts)

    keeps = [kept_rows(wl, l) - w for l in range(nl)]
    from kvcache_factory_b200 import pyramidkv_utils as _U
    use_half = clusters[0].async_scores and groups * w <= 128 and w <= 32 and d == 128

    qkw = dict(seq_len=L) if q_window_only else {}

    def scores_phase():
        """The scoring launches exactly as the step issues them: with async_scores the half-footprint kernel on
        the rotating high-priority streams (two launches share every SM), else the full kernel on one stream."""
        if not use_half:
            return [_native.window_scores(Qs[l], Ks[l], w, "sum", impl, **qkw) for l in range(nl)]
        cur, outs, used = torch.cuda.current_stream(), [], []
        for l in range(nl):
            st = _U.compression_stream(dev, 100 + l % _U.NUM_SCORE_STREAMS)
            if st not in used:
                st.wait_stream(cur); used.append(st)
            with torch.cuda.stream(st):
                outs.append(_native.window_scores(Qs[l], Ks[l], w, "sum", _native.IMPL_TCGEN05_HALF, **qkw))
        for st in used:
            cur.wait_stream(st)
        return outs

    g_sc, scs = phase_graph(scores_phase)
    g_solo, _solo = phase_graph(lambda: [_native.window_scores(Qs[l], Ks[l], w, "sum", impl, **qkw) for l in range(nl)])
    g_se, idxs = phase_graph(lambda: [_native.pool_topk(scs[l], wl["kernel"], wl["pooling"], keeps[l])[0]
                                      for l in range(nl)])
    g_ga, _outs = phase_graph(lambda: [_native.gather_dense(Ks[l], Vs[l], idxs[l], Hq, keeps[l], w) for l in range(nl)])
    reps = max(3, min(args.steps, 10))
    tot = {"scores": time_graph(g_sc, reps), "select": time_graph(g_se, reps), "gather": time_graph(g_ga, reps)}
    solo_scores_ms = time_graph(g_solo, reps)
    del g_sc, g_solo, _solo, g_se, g_ga, scs, idxs, _outs
    dominant = max(tot, key=tot.get)
    sb = sum(alg_bytes_layer(wl, l)[0] for l in range(nl))
    gb = sum(alg_bytes_layer(wl, l)[1] for l in range(nl))
    alg = {"scores": sb, "select": 0, "gather": gb}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_tf = float(peaks.get("bf16_tflops_sustained", 0.0)) or None   # H2O runs for ~0.2 s per layer: the sustained figure
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom_bytes_per_launch = alg[dominant] / nl if dominant != "select" else (sb + gb) / nl
    dom_ms = tot[dominant] / nl
    achieved = dom_bytes_per_launch / (dom_ms * 1e-3) / 1e9
    # DRAM traffic per launch of the dominant kernel from the committed ncu --set full capture
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2b_traffic.json")))
        kname = {"scores": "window_score_half_kernel" if use_half else "window_score_fused_kernel",
                 "select": "topk_cluster_kernel", "gather": "gather_rows_kernel"}
        if args.workload == "pyramidkv_llama3_8b_32k":
            traffic = tj["dram_bytes_per_launch"].get(kname[dominant])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": {"scores": ("window_scores (fused stats+scores passes; half-footprint CTAs, two launches in flight "
                                                      "per SM: ms_per_launch = phase time / launches)" if use_half else
                                                      "window_scores (fused stats+scores passes, one launch)"),
                                           "select": "pool_topk", "gather": "gather_rows"}[dominant],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms,
                "phase_ms_per_step": tot, "scores_full_kernel_back_to_back_ms_per_launch": solo_scores_ms / nl,
                "step_frac_of_hbm_roofline": (sb + gb) / (ms_per_step * 1e-3) / 1e9 / peak}

    # ---- e2e: host buffers -> compressed host buffers --------------------------------------------
    e2e = None
    if not args.no_e2e:
        hc = HostCompressor(clusters, groups, dev)
        hK = [k.cpu().pin_memory() for k in Ks]
        hV = [v.cpu().pin_memory() for v in Vs]
        hQ = [(q if q_window_only else q[:, :, L - w:, :]).contiguous().cpu().pin_memory() for q in Qs]   # the only rows the path reads
        hc.run(hK, hQ, hV, L)                                                 # warm-up (allocates pinned outputs)
        hc.run(hK, hQ, hV, L)
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_e2e):
            hc.run(hK, hQ, hV, L)
        t1.record()
        barrier()
        tt = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / n_e2e
        e2e = {"value": world * B * L / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": hc.h2d_bytes, "d2h_bytes_per_step": hc.d2h_bytes, "steps": n_e2e}
        del hK, hV, hQ

    # ---- side blocks: free the headline inputs first ----
    del Ks, Vs, Qs, clusters
    if graph is not None:
        del graph, graph_out
    torch.cuda.empty_cache()
    cfg5 = cfg4 = cfg1 = cfg3 = checks = None
    if not args.no_extra and args.workload == "pyramidkv_llama3_8b_32k":
        cfg5 = measure_cfg5(torch, dist if world > 1 else None, dev, rank, world, args.steps, not args.sync_tail,
                            not args.sync_tail and not args.sync_scores, peak)
        if world == 1:
            cfg4 = measure_cfg4(torch, dev, peak)
            cfg1 = measure_cfg1(torch, dev, peak)
            torch.cuda.empty_cache()
            cfg3 = measure_cfg3(torch, dev, peak, peak_tf)
        else:
            checks = check_sharded_adakv(torch, dist, dev, rank, world)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            n = wl["layers"]
            sample = sorted(set([0, n // 2, n - 1]))
            v, per_l, cores, kind = cpu_reference(wl, sample)
            src = "the unmodified reference file (oracle/_ref copy)" if kind == "reference" else "oracle/kvc_oracle.py"
            cpu = {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind,
                   "sample": f"{src}: repeat_kv + update_kv on layers {sample} of {n} at full L={L} "
                             f"({per_l * 1e3:.0f} ms/layer) x {n} layers"}
        line = {
            "metric": "kv_tokens_compressed_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": config_of(args.workload, wl, world),
            "impl_flags": {"score_impl": args.score_impl, "cuda_graph": used_graph, "async_tail": not args.sync_tail,
                           "async_scores": not args.sync_tail and not args.sync_scores,
                           "note": "value = one CUDA-graph replay per step with the asynchronous select/gather tails and "
                                   "two scoring launches in flight (config.kvc_async_tail / kvc_async_scores through the "
                                   "hooks); sync_ms_per_step = the drop-in defaults"},
            "sync_ms_per_step": sync_ms,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "cfg5": cfg5, "cfg4": cfg4, "cfg1": cfg1, "cfg3": cfg3, "checks": checks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pyramidkv_llama3_8b_32k", choices=sorted(WORKLOADS))
    ap.add_argument("--score-impl", default="auto", choices=["auto", "mma", "tc", "fused", "half", "dyn"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--sync-tail", action="store_true", help="keep select+gather on the caller's stream")
    ap.add_argument("--sync-scores", action="store_true", help="keep the scoring kernel on the caller's stream")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg5 / cfg4 / NCCL-check side blocks")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
