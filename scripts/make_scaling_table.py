"""profiles/r2_scaling.txt from gpurun_out/scaling_{1,2,4,8}.txt (written on the GPU box by scripts/scaling_r2.sh N)."""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
b, ceil = {}, {}
for n in (1, 2, 4, 8):
    path = os.path.join(root, "gpurun_out", f"scaling_{n}.txt")
    if not os.path.exists(path):
        continue
    for line in open(path):
        if line.startswith('{"metric"'):
            b[n] = json.loads(line)
        elif line.startswith('{"h2d'):
            ceil[n] = json.loads(line)
ns = sorted(b)
out = ["round 2 -- scaling on one 8xB200 box (bench.py --gpus N --steps 5 --warmup 3 under torchrun; scripts/scaling_r2.sh N)", "",
       "weak scaling (value: N cfg-2 requests, 8 (batch,kv-head) units per rank, no collective):",
       "  N   ms/step   tokens/s (all ranks)   vs N=1 x N"]
v1 = b[ns[0]]["value"] / ns[0]
for n in ns:
    out.append(f"  {n}   {b[n]['ms_per_step']:.3f}     {b[n]['value'] / 1e6:8.1f} M            {b[n]['value'] / (v1 * n):.3f}")
out += ["", "cfg5 = BASELINE configs[4] (Llama-3-70B, one 131072-token prefill, 80 layers, 8 KV heads sharded 8/N per rank, HBM-cold;",
        "       strong scaling, N=1 re-measured in the same run on rank 0 as the denominator):",
        "  N   heads/rank   ms/step   n1 ms/step (same box)   efficiency   hbm frac per rank   us per layer"]
for n in ns:
    c = b[n]["cfg5"]
    out.append(f"  {n}   {c['kv_heads_per_rank']}            {c['ms_per_step']:6.2f}    {c.get('n1_ms_per_step_same_box', c['ms_per_step']):6.2f}"
               f"                  {c.get('strong_scaling_efficiency', 1.0):.3f}        {c['hbm_frac_per_rank']:.3f}               {c['ms_per_step'] / 80 * 1e3:5.1f}")
out += ["", "  every 128-row block of a KV head is a unit of the half-footprint scoring kernel (two launches in flight): one head at",
        "  128K is 27.5 us per layer of scoring (41.0 us with the two-accumulator full-size kernel, which ran alone), the select of",
        "  8 rows x 131040 (19 us under ncu) and the gather (7 us) run on side streams between the scoring CTAs.", "",
        "e2e leg (host buffers, every rank moves 2.70 GB in and 1.08 GB out per step) against the box's plain cudaMemcpyAsync ceiling",
        "(scripts/h2d_ceiling.py: all ranks copy pinned 1 GiB buffers at once; GB/s per GPU per direction):",
        "  N   e2e ms/step   e2e tokens/s   in GB/s per GPU achieved   ceiling h2d-only   d2h-only   both at once"]
for n in ns:
    e, c = b[n]["e2e"], ceil.get(n)
    ach = e["h2d_bytes_per_step"] / e["ms_per_step"] / 1e6
    cs = (f"{c['h2d_GBps_per_gpu_per_direction']:6.1f}            {c['d2h_GBps_per_gpu_per_direction']:6.1f}     {c['both_GBps_per_gpu_per_direction']:6.1f}"
          if c else "   -")
    out.append(f"  {n}   {e['ms_per_step']:7.1f}       {e['value'] / 1e3:7.0f} k      {ach:6.1f}                     {cs}")
out += ["", "  the host side of the box delivers ~185-210 GB/s h2d in total at N=4 and N=8 alike, and much less with both directions",
        "  active: the e2e leg is bound by the host's PCIe / memory system, not by the GPUs (at N=1 it runs at 51.5 of the 53-55 GB/s",
        "  a plain copy reaches).", "",
        "NCCL checks at every N>1 (bench 'checks' block): sharded AdaKV (two-level histogram exchange) capacities and rows equal the",
        "single-GPU cluster's: " + ", ".join(f"N={n}: {b[n]['checks']['sharded_adakv_capacities_equal_single_gpu'] and b[n]['checks']['sharded_adakv_rows_equal_single_gpu']}" for n in ns if b[n].get("checks"))]
open(os.path.join(root, "profiles", "r2_scaling.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
