import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, mvgeo
from oracle import mvgeo_oracle as O
DEV="cuda:0"
n_maps,H,W,dtype=512,480,640,torch.bfloat16
g = torch.Generator(device=DEV); g.manual_seed(31)
kp = torch.rand((n_maps, 2), generator=g, device=DEV) * torch.tensor([W - 1.0, H - 1.0], device=DEV)
maps = mvgeo.encode_gaussian(kp, (H, W), 3.0, dtype)
amp = torch.rand((n_maps, 1, 1), generator=g, device=DEV) * 0.95 + 0.05
step = max(1, n_maps // 16)
for m0 in range(0, n_maps, step):
    sl = slice(m0, m0 + step)
    noise = torch.randn(maps[sl].shape, generator=g, device=DEV) * 0.01
    maps[sl] = (maps[sl].float() * amp[sl] + noise).to(dtype)
maps[::7] = torch.rand(maps[::7].shape, generator=g, device=DEV).to(dtype)
maps[::11] = (torch.round(maps[::11].float() * 8) / 8).to(dtype)
r = mvgeo.decode_heatmaps(maps, None, soft="global", beta=25.0)
seen = maps.float().cpu().numpy()
sub = np.arange(0, n_maps, 5)
ref = O.soft_argmax(seen[sub], 25.0, "global")
got = r.kp_soft.cpu().numpy()[sub]
err = np.abs(got-ref).max(axis=1)
order = np.argsort(-err)[:12]
for o in order:
    m = sub[o]
    kind = "uniform" if m % 7 == 0 else ("quant" if m % 11 == 0 else "blob")
    print(m, kind, "amp", float(amp[m]), "err", err[o], "peak", seen[m].max(), "got", got[o], "ref", ref[o])
# emulate: f32 weights, f64 sums -> isolates accumulation from exp error
m = sub[order[0]]
h = seen[m].astype(np.float64)
w32 = np.exp2((seen[m].astype(np.float32) - np.float32(seen[m].max())) * np.float32(25.0*1.4426950408889634)).astype(np.float64)
ys, xs = np.mgrid[0:H,0:W]
print("f32-weight f64-sum centroid", (w32*xs).sum()/w32.sum(), (w32*ys).sum()/w32.sum())
