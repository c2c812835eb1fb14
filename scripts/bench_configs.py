"""Per-layer device time of every BASELINE.json configuration that is not the bench.py headline
(cfg-1, cfg-3 StreamingLLM/H2O, cfg-4 AdaKV prefill + flattened decode append, cfg-5 one 70B shard).
One JSON line per configuration; CUDA events, K/V rotate over buffers larger than L2 where they fit.
usage: python scripts/bench_configs.py [cfg ...]   (default: all)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import (AdaKVCluster, DynamicCacheSplitHeadFlatten, H2OKVCluster, PyramidKVCluster,
                                  SnapKVCluster, StreamingLLMKVCluster, _native)

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1370.8}


def rand(*shape, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32).to(dtype)


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def timed_graph(fn, reps=8, inner=8):
    """Same call captured `inner` times into one CUDA graph: launch overhead of the Python wrapper excluded."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            keep = [fn() for _ in range(inner)]
    torch.cuda.current_stream().wait_stream(cap)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / inner)
    ts.sort()
    del keep
    return ts[len(ts) // 2], ts[0]


def emit(name, us_med, us_best, tokens, alg_bytes=None, alg_flops=None, **extra):
    line = {"config": name, "us_per_layer_median": round(us_med, 1), "us_per_layer_best": round(us_best, 1),
            "tokens_per_s_per_layer": round(tokens / (us_med * 1e-6)), **extra}
    if alg_bytes:
        line["alg_MB"] = round(alg_bytes / 1e6, 1)
        line["hbm_frac"] = round(alg_bytes / (us_med * 1e-6) / 1e9 / PEAK["hbm_gbs"], 3)
    if alg_flops:
        line["tflops"] = round(alg_flops / (us_med * 1e-6) / 1e12, 1)
        line["tensor_frac"] = round(line["tflops"] / PEAK.get("bf16_tflops_sustained", 1370.8), 3)
    print(json.dumps(line), flush=True)


def cfg1():
    B, Hq, Hkv, L, d, w, cap = 1, 32, 8, 4096, 128, 32, 128
    K, V, Q = rand(B, Hkv, L, d, seed=1), rand(B, Hkv, L, d, seed=2), rand(B, Hq, L, d, seed=3)
    for dt in (torch.bfloat16, torch.float16):
        cl = SnapKVCluster(w, cap, 7, "avgpool")
        k, v, q = K.to(dt), V.to(dt), Q.to(dt)
        alg = B * Hkv * L * d * 2 + B * Hq * w * d * 2 + 4 * B * Hq * cap * d * 2
        med, best = timed(lambda: cl.update_kv(k, q, v, None, 4), iters=20)
        emit(f"cfg1 SnapKV 8B L=4096 w=32 cap=128 avgpool7 {str(dt)[6:]} (3 launches, eager)", med, best, B * L, alg)
        med, best = timed_graph(lambda: cl.update_kv(k, q, v, None, 4))
        emit(f"cfg1 SnapKV 8B L=4096 w=32 cap=128 avgpool7 {str(dt)[6:]} (3 launches, CUDA graph)", med, best, B * L, alg)


def cfg3():
    B, Hq, Hkv, L, d, w, cap = 8, 32, 8, 32768, 128, 32, 1024
    K, V = rand(B, Hkv, L, d, seed=1), rand(B, Hkv, L, d, seed=2)
    Q = rand(B, Hq, L, d, seed=3)
    cl = StreamingLLMKVCluster(w, cap, 7, "avgpool")
    med, best = timed(lambda: cl.update_kv(K, Q, V, None, 4), iters=10)
    emit("cfg3 StreamingLLM Mistral-7B B=8 L=32768 cap=1024 (eager)", med, best, B * L, 4 * B * Hq * cap * d * 2)
    med, best = timed_graph(lambda: cl.update_kv(K, Q, V, None, 4))
    emit("cfg3 StreamingLLM Mistral-7B B=8 L=32768 cap=1024 (CUDA graph)", med, best, B * L, 4 * B * Hq * cap * d * 2)
    cl = H2OKVCluster(w, cap, 7, "avgpool")
    med, best = timed(lambda: cl.update_kv(K, Q, V, None, 4), iters=2, warm=1)
    emit("cfg3 H2O Mistral-7B B=8 L=32768 cap=1024 (all-query scores, two tcgen05 passes)", med, best, B * L,
         alg_flops=4.0 * B * Hq * L * L * d)


def cfg4():
    B, Hq, Hkv, L, d, w, cap = 1, 32, 8, 65536, 128, 32, 2048
    K, V, Q = rand(B, Hkv, L, d, seed=1), rand(B, Hkv, L, d, seed=2), rand(B, Hq, L, d, seed=3)
    Kr, Vr = K.repeat_interleave(4, 1), V.repeat_interleave(4, 1)   # the reference's AdaKV call sites pass repeated K/V
    out = {}
    for label, kk, vv in (("repeat_kv-expanded K/V as the reference passes them", Kr, Vr),
                          ("un-repeated K/V (GQA folded in the kernels)", K, V)):
        cl = AdaKVCluster(window_size=w, kernel_size=7, pooling="maxpool", max_capacity_prompt=cap, floor=0.2,
                          normalize=True, layer_idx=0, num_hidden_layers=32)
        def run():
            out["kv"] = cl.update_kv(kk, Q, vv)
        med, best = timed(run, iters=5)
        rows = out["kv"][0].shape[0]
        alg = Hkv * L * d * 2 + Hq * w * d * 2 + 4 * rows * d * 2
        emit(f"cfg4 AdaKV 8B L=65536 cap=2048 floor=0.2 maxpool7 prefill update_kv, {label} (incl. allocation + the "
             "reference's host read of the capacities)", med, best, L, alg, flat_rows=int(rows))
    # decode: flattened append of one token per head (cuda_api.cu update_flatten_view)
    flatK = out["kv"][0]
    state = rand(Hq, d, seed=9)
    head_lens, cu = cl.head_lens.clone(), cl.cu_klen.clone()
    med, best = timed(lambda: _native.update_flatten_view(flatK, state, head_lens, cu), iters=20)
    emit("cfg4 AdaKV decode update_flatten_view (one tensor, one layer-step, eager)", med, best, 1,
         (flatK.shape[0] * 2 + Hq) * d * 2)
    med, best = timed_graph(lambda: _native.update_flatten_view(flatK, state, head_lens, cu))
    emit("cfg4 AdaKV decode update_flatten_view (one tensor, one layer-step, CUDA graph)", med, best, 1,
         (flatK.shape[0] * 2 + Hq) * d * 2)


def cfg5():
    # one GPU's share at 8 GPUs: 1 KV head (8 query heads) of Llama-3-70B at 128K
    B, Hq, Hkv, L, d, w, cap = 1, 8, 1, 131072, 128, 32, 2048
    Ks = [rand(B, Hkv, L, d, seed=i) for i in range(6)]   # 6 x 32 MiB of K rotate (> L2 with V)
    Vs = [rand(B, Hkv, L, d, seed=10 + i) for i in range(6)]
    Qw = rand(B, Hq, w, d, seed=3)
    for layer in (0, 79):
        cl = PyramidKVCluster(num_hidden_layers=80, window_size=w, max_capacity_prompt=cap, kernel_size=7,
                              pooling="avgpool", layer_idx=layer)
        it = [0]
        res = {}
        def run():
            i = it[0] % 6; it[0] += 1
            res["o"] = cl.update_kv_window(Ks[i], Qw, Vs[i], 8, L)
        med, best = timed(run, iters=12, warm=3)
        C = res["o"][0].shape[2]
        alg = Hkv * L * d * 2 + Hq * w * d * 2 + 4 * Hq * C * d * 2
        emit(f"cfg5 PyramidKV 70B shard (1 KV head / 8 q heads) L=131072 layer {layer} kept {C} (eager)", med, best, L, alg)
        def run_g():
            i = it[0] % 6; it[0] += 1
            return cl.update_kv_window(Ks[i], Qw, Vs[i], 8, L)
        med, best = timed_graph(run_g, inner=12)
        emit(f"cfg5 PyramidKV 70B shard (1 KV head / 8 q heads) L=131072 layer {layer} kept {C} (CUDA graph)", med, best, L, alg)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]
    for c in todo:
        try:
            globals()[c]()
        except Exception as e:  # keep going: one line per failure
            print(json.dumps({"config": c, "error": f"{type(e).__name__}: {e}"[:400]}), flush=True)
        torch.cuda.empty_cache()
