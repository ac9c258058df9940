"""Small exercise of every decode-kernel variant for compute-sanitizer (memcheck / racecheck):
all dtypes x modes x consumer-group counts, multi-map walking with ring wrap, partial last tiles, per-view
pointers, the fused loss pass, the generic kernel, and the PnP solver."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvgeo
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
shapes = [(700, 16, 64), (610, 128, 128), (330, 240, 320), (40, 480, 640), (9, 37, 41), (5, 24, 40)]
for n, H, W in shapes:
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        if H * W * n > 3e7 and dt == torch.float32:
            n_ = n // 4
        else:
            n_ = n
        m = torch.randn((n_, H, W), generator=g, device=dev).to(dt)
        for soft in (None, "global", "window"):
            r = mvgeo.decode_heatmaps(m, (1200, 1920), soft=soft, beta=30.0, window_radius=3)
        if (W * m.element_size()) % 64 == 0 and (H * W * m.element_size()) % 128 == 0:
            kp = torch.rand((n_, 2), generator=g, device=dev) * torch.tensor([W - 1.0, H - 1.0], device=dev)
            loss, dec = mvgeo.decode_and_mse(m, kp, 3.0, 10.0)
views = [torch.randn((6, 7, 64, 96), generator=g, device=dev).to(torch.bfloat16) for _ in range(3)]
mvgeo.decode_heatmaps(views, None, soft="global", beta=10.0)
chain = mvgeo.Chain.builtin("fr3")
q = torch.rand((16, 7), generator=g, device=dev) * 2 - 1
rig = mvgeo.CameraRig.synthetic_ring(2, distortion=True)
X = mvgeo.forward_kinematics(chain, q)[:, 0]
kp = mvgeo.project_points(X, rig)
mvgeo.pnp_solve(X, kp, rig)
mvgeo.pnp_refine(X, kp, rig)
torch.cuda.synchronize()
print("sanitize run complete")
