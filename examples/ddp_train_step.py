#!/usr/bin/env python
"""Config 4 harness: Fr5 DDP training step around the hot path's two loss kernels.

    python examples/ddp_train_step.py --steps 20                       # 1 GPU
    torchrun --nproc-per-node 8 examples/ddp_train_step.py --steps 20  # 8 GPUs, NCCL grad all-reduce

What the reference does per step (model/MvRoPose_FR3.py:783-861, model/Fr5_model_train.ipynb):
frozen backbone -> trainable heads -> nn.MSELoss(heat-maps, GT maps) * weight + angle loss ->
backward -> DDP bucketed all-reduce -> AdamW. GT maps are rasterised on the CPU by DataLoader
workers (FK -> project -> create_gt_heatmap, MvRoPose_FR3.py:214-222) and copied to the GPU.

Here the heads are a small stand-in (the backbone is out of scope: it stays on cuDNN/cuBLAS), and
everything between "angles / heat-maps out of the network" and "scalar loss" runs in the mvgeo
kernels, on the device, with gradients:
  * GT key-points = project(FK(gt_angles))                          (mvgeo_fk + mvgeo_project)
  * L_kpt = heatmap_mse_loss(pred_maps, gt_kp, sigma=5) * 1e4       (mvgeo_heatmap_mse fwd+bwd, no GT maps materialised)
  * L_fk  = fk_reproj_loss(pred_angles, gt_uv)                      (mvgeo_fk_reproj_fwd/bwd: d loss / d angles)
DDP's gradient all-reduce over NCCL is untouched. Frames shard across ranks (DistributedSampler-style).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

import mvgeo


class Heads(nn.Module):
    """Stand-in for UNetViTKeypointHead + JointAngleHead (model/MvRoPose_FR3.py:355,516): feature maps
    (B*V, C, 32, 32) -> heat-maps (B*V, K, 128, 128) and joint angles (B, J) in degrees."""

    def __init__(self, C, K, J, V):
        super().__init__()
        self.V = V
        self.kpt = nn.Sequential(nn.Conv2d(C, 64, 3, padding=1), nn.GELU(), nn.Upsample(scale_factor=4, mode="bilinear"),
                                 nn.Conv2d(64, K, 3, padding=1))
        self.ang = nn.Sequential(nn.Linear(C * V, 256), nn.GELU(), nn.Linear(256, J))

    def forward(self, feats):
        maps = self.kpt(feats)
        pooled = feats.mean(dim=(2, 3)).reshape(-1, self.V * feats.shape[1])
        return maps, self.ang(pooled) * 90.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    robot, V, C, HM = "fr5", 3, 96, 128
    chain = mvgeo.Chain.builtin(robot)
    K, J = chain.n_points, chain.n_joints
    rig = mvgeo.CameraRig.synthetic_ring_for(robot, V, distortion=True)
    Rv = np.stack([np.asarray(mvgeo.view_rotation(robot, v)) for v in ("top", "left", "right")]).astype(np.float32)
    Hi, Wi = rig.image_size
    torch.manual_seed(1234)  # identical initial weights on every rank
    model = Heads(C, K, J, V).to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)  # each rank owns its own frames
    B = args.batch
    to_map = torch.tensor([HM / Wi, HM / Hi], device=dev)

    def step():
        feats = torch.randn((B * V, C, 32, 32), generator=g, device=dev)
        gt_q = (torch.rand((B, J), generator=g, device=dev) * 2 - 1) * 120.0          # degrees
        with torch.no_grad():                                                         # GT on the device
            gt_uv = mvgeo.project_points(mvgeo.forward_kinematics(chain, gt_q, Rv), rig)   # (B,V,K,2) image px
            gt_kp = (gt_uv * to_map).reshape(B * V, K, 2)                                   # map px
        maps, q = model(feats)
        l_kpt = mvgeo.heatmap_mse_loss(maps, gt_kp, sigma=5.0, weight=1e4)
        l_fk, _, _, _ = mvgeo.fk_reproj_loss(chain, q.float(), rig, gt_uv, Rv, lam=1e-4)
        l_ang = nn.functional.smooth_l1_loss(q, gt_q)
        loss = l_kpt + l_fk + l_ang
        opt.zero_grad(set_to_none=True)
        loss.backward()                                                               # DDP all-reduce here
        opt.step()
        return loss.detach(), l_kpt.detach(), l_fk.detach()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    hist = [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    first, last = [float(x) for x in hist[0]], [float(x) for x in hist[-1]]
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    if rank == 0:
        print(json.dumps({"config": "C4: Fr5 DDP training step (heat-map MSE + DH reprojection loss fwd/bwd, grad all-reduce)",
                          "n_gpus": world, "frames_per_gpu": B, "views": V, "steps": args.steps,
                          "ms_per_step": 1e3 * dt / args.steps, "frames_per_s": B * world * args.steps / dt,
                          "loss_first": first, "loss_last": last}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
