"""Kernel-development micro-benchmark: decode kernel alone, CUDA-event timed.
usage: python tools/perf_decode.py [B V K H W dtype mode]   (defaults: C2 shape, bf16, global)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvgeo

a = sys.argv[1:]
B, V, K, H, W = (int(x) for x in a[:5]) if len(a) >= 5 else (1024, 4, 8, 240, 320)
dtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a[5] if len(a) > 5 else "bf16"]
mode = a[6] if len(a) > 6 else "global"
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(1)
kp = torch.rand((B, V, K, 2), generator=g, device=dev) * torch.tensor([W - 1.0, H - 1.0], device=dev)
maps = mvgeo.encode_gaussian(kp, (H, W), 3.0, dtype)
for b0 in range(0, B, 64):
    sl = maps[b0:b0 + 64]
    sl.add_(torch.randn(sl.shape, generator=g, device=dev, dtype=torch.float32).mul_(0.01).to(dtype))
nbytes = maps.numel() * maps.element_size()
import ctypes as C
lib = mvgeo._lib.load()
n_maps = B * V * K
idx = torch.empty((n_maps,), dtype=torch.int32, device=dev)
peak, score = torch.empty((n_maps,), device=dev), torch.empty((n_maps,), device=dev)
kph, kps = torch.empty((n_maps, 2), device=dev), torch.empty((n_maps, 2), device=dev)
DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
MODE = {"none": 0, "global": 1, "window": 2}[mode]
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.mvgeo_decode(maps.data_ptr(), DT, n_maps, H, W, 1920 / W, 1200 / H, MODE, 100.0, 3, 0, 1, 1, 0, idx.data_ptr(),
                          peak.data_ptr(), score.data_ptr(), kph.data_ptr(), kps.data_ptr(), st)
    assert rc == 0, rc
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
med, best = statistics.median(ts), min(ts)
chk = int(idx.sum()), float(kps.double().sum())
tag = "libmvgeo"
print(f"{tag:28s} {B}x{V}x{K}x{H}x{W} {a[5] if len(a)>5 else 'bf16'} {mode:6s} "
      f"median {med*1e3:8.1f} us  {nbytes/med/1e6:8.1f} GB/s   best {nbytes/best/1e6:8.1f} GB/s  chk={chk}")
