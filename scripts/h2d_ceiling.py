"""Plain concurrent cudaMemcpyAsync ceilings of the box (no kernels of ours): every rank copies a pinned 1 GiB buffer
host->device, device->host and both at once, all ranks at the same time.  Explains the e2e leg of bench.py at N > 1.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_ceiling.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_dir = reps * n / float(t.item()) / 1e9
    return per_dir


res = {}
for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    res[mode + "_GBps_per_gpu_per_direction"] = round(run(mode), 2)
if rank == 0:
    numa = "?"
    try:
        numa = open(f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else ''}/numa_node").read().strip()
    except Exception:
        pass
    res.update(n_gpus=world, host_cores=os.cpu_count())
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
