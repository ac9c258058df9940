"""Host-to-device ceiling of the box: every rank copies a pinned host buffer to its GPU at the same time.
    torchrun --nproc-per-node N tools/h2d_probe.py [--numa]      (or plain python for N = 1)
Prints one JSON line: per-rank and aggregate GB/s, with and without binding each rank to its GPU's CPUs
before the pinned allocation (first touch). Names the limiter of the e2e leg of bench.py at N > 1."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mvgeo

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=2.0)
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(a.gb * 1e9)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
out = {}
for mode in ("unbound", "numa_bound"):
    cpus = None
    if mode == "numa_bound":
        cpus = mvgeo.sharding.bind_to_gpu_numa(local)
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    for direction in ("h2d", "d2h"):
        for _ in range(2):
            (dst.copy_(src, non_blocking=True) if direction == "h2d" else src.copy_(dst, non_blocking=True))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            (dst.copy_(src, non_blocking=True) if direction == "h2d" else src.copy_(dst, non_blocking=True))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([n * a.reps / dt / 1e9], device=dev, dtype=torch.float64)
        if world > 1:
            allv = [torch.zeros_like(gbs) for _ in range(world)]
            dist.all_gather(allv, gbs)
            vals = [float(v) for v in allv]
        else:
            vals = [float(gbs)]
        out[f"{mode}_{direction}"] = {"per_rank_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(sum(vals), 1)}
    out[f"{mode}_cpus"] = len(cpus) if cpus else 0
    del src
if rank == 0:
    print(json.dumps({"probe": "concurrent pinned-host <-> device copies", "n_gpus": world, "gb_per_copy": a.gb, **out}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
