"""Kernel-development probe: all_gather_into_tensor latency vs message size at N ranks."""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ws = dist.get_world_size()
for mb in (1, 4, 16, 18, 64, 128):
    n = mb * 1024 * 1024 // 4
    src = torch.ones(n, device=dev); dst = torch.empty(n * ws, device=dev)
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather_into_tensor(dst, src); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if rank == 0:
        print(f"ws={ws} {mb:4d} MB/rank: " + " ".join(f"{t:7.3f}" for t in ts) + f" ms   busbw(last) {mb*1.048576e-3*(ws-1)/ts[-1]*1e3:7.1f} GB/s", flush=True)
dist.destroy_process_group()
