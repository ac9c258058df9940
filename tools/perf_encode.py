"""Kernel-development micro-benchmark: GT encoder and fused heat-map MSE (fwd, fwd+bwd)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvgeo

a = sys.argv[1:]
B, V, K, H, W = (int(x) for x in a[:5]) if len(a) >= 5 else (1024, 4, 8, 240, 320)
dtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a[5] if len(a) > 5 else "bf16"]
dev = "cuda:0"
lib = mvgeo._lib.load()
n_maps = B * V * K
g = torch.Generator(device=dev); g.manual_seed(1)
kp = torch.rand((n_maps, 2), generator=g, device=dev) * torch.tensor([W - 1.0, H - 1.0], device=dev)
maps = torch.empty((n_maps, H, W), dtype=dtype, device=dev)
grad = torch.empty_like(maps)
partial = torch.empty((n_maps,), device=dev); loss = torch.empty((), device=dev)
DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
st = torch.cuda.current_stream().cuda_stream
nbytes = maps.numel() * maps.element_size()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
enc = lambda: lib.mvgeo_encode_gaussian(kp.data_ptr(), n_maps, H, W, 3.0, DT, maps.data_ptr(), st)
fwd = lambda: lib.mvgeo_heatmap_mse(maps.data_ptr(), DT, kp.data_ptr(), n_maps, H, W, 3.0, 100.0, None, partial.data_ptr(), loss.data_ptr(), None, st)
bwd = lambda: lib.mvgeo_heatmap_mse(maps.data_ptr(), DT, kp.data_ptr(), n_maps, H, W, 3.0, 100.0, None, partial.data_ptr(), loss.data_ptr(), grad.data_ptr(), st)
te, tf, tb = t(enc), t(fwd), t(bwd)
print(f"{B}x{V}x{K}x{H}x{W} {a[5] if len(a)>5 else 'bf16'}: encode {te*1e3:7.1f} us {nbytes/te/1e6:7.1f} GB/s (write) | mse fwd {tf*1e3:7.1f} us {nbytes/tf/1e6:7.1f} GB/s (read)"
      f" | mse fwd+bwd {tb*1e3:7.1f} us {2*nbytes/tb/1e6:7.1f} GB/s (read+write)  loss={float(loss):.4g}")
