#!/usr/bin/env python
"""Config 4 harness: Fr5 DDP training step around the hot path's loss kernels, with a per-phase breakdown.

    python examples/ddp_train_step.py --steps 20                       # 1 GPU
    torchrun --nproc-per-node 8 examples/ddp_train_step.py --steps 20  # 8 GPUs, NCCL grad all-reduce

What the reference does per step (model/MvRoPose_FR3.py:783-861, model/Fr5_model_train.ipynb:4428-4438:
batch 150, heat-map weight 1e4): frozen backbone (requires_grad=False, model/DREAM_Train.py:739-740) ->
trainable heads -> nn.MSELoss(heat-maps, GT maps) * weight + angle loss -> ONE backward -> DDP bucketed
all-reduce of the trainable parameters (model/MvRoPose_FR3.py:973) -> AdamW. GT maps are rasterised on the
CPU by DataLoader workers (FK -> project -> create_gt_heatmap, MvRoPose_FR3.py:214-222) and copied to the GPU.

Here the backbone is out of scope (it stays on cuDNN / cuBLAS): a frozen stand-in produces the feature maps
under no_grad, the trainable heads are a stand-in sized like the reference's (--head-params, default ~25 M), and
everything between "angles / heat-maps out of the network" and "scalar loss" runs in the mvgeo kernels, on the
device, with gradients:
  * GT key-points = project(FK(gt_angles))                          (mvgeo_fk + mvgeo_project)
  * L_kpt = heatmap_mse_loss(pred_maps, gt_kp, sigma=5) * 1e4       (mvgeo_heatmap_mse fwd+bwd, no GT maps materialised)
    and, with --kp-metric fused (default), the per-step key-point pixel error of the predicted maps from the SAME read:
    mvgeo.decode_and_mse = mvgeo_decode_mse (loss + arg-max in one pass over the maps)
  * L_fk  = fk_reproj_loss(pred_angles, gt_uv)                      (mvgeo_fk_reproj_fwd/bwd: d loss / d angles)
DDP's gradient all-reduce over NCCL is untouched: only the trainable heads are wrapped, so only their
gradients are bucketed. The breakdown (CUDA events, max over ranks) names the all-reduce share two ways:
the step with and without gradient synchronisation (model.no_sync()), and a stand-alone all-reduce of the
same number of bytes.
"""
import argparse
import contextlib
import json
import os
import statistics
import sys

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

    def __init__(self, C, K, J, V, width):
        super().__init__()
        self.V = V
        # (the 4x up-sampling acts on the K heat-map channels: torch's bilinear backward launches one grid row per
        # (image, channel) and refuses 450 x ~1000 of them)
        self.kpt = nn.Sequential(nn.Conv2d(C, width, 3, padding=1), nn.GELU(), nn.Conv2d(width, width, 3, padding=1), nn.GELU(),
                                 nn.Conv2d(width, K, 3, padding=1), nn.Upsample(scale_factor=4, mode="bilinear"))
        self.ang = nn.Sequential(nn.Linear(C * V, 4 * width), nn.GELU(), nn.Linear(4 * width, 4 * width), nn.GELU(),
                                 nn.Linear(4 * width, J))

    def forward(self, feats):
        maps = self.kpt(feats)
        pooled = feats.mean(dim=(2, 3)).reshape(-1, self.V * feats.shape[1])
        return maps, self.ang(pooled) * 90.0


def width_for(target_params, C, K, V):
    """Head width whose parameter count is closest to target_params (conv 9*C*w + 9*w*w + 9*w*K, MLP 4w*(CV + 4w))."""
    best, best_w = None, 64
    for w in range(32, 4096, 16):
        n = 9 * C * w + 9 * w * w + 9 * w * K + 4 * w * (C * V) + 16 * w * w
        if best is None or abs(n - target_params) < best:
            best, best_w = abs(n - target_params), w
    return best_w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=150, help="frames per GPU (reference: 150, Fr5_model_train.ipynb:4428-4438)")
    ap.add_argument("--head-params", type=float, default=25e6, help="trainable parameters of the stand-in heads")
    ap.add_argument("--bf16", action="store_true", help="heads under bf16 autocast (heat-maps reach the loss kernels in bf16)")
    ap.add_argument("--kp-metric", choices=("fused", "separate", "none"), default="fused",
                    help="key-point pixel error of the predicted maps per step: 'fused' = mvgeo.decode_and_mse (loss and "
                         "arg-max from ONE read of the maps), 'separate' = heatmap_mse_loss + decode_heatmaps (two reads)")
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
    backbone = nn.Sequential(nn.Conv2d(3, C, 7, stride=4, padding=3), nn.GELU(), nn.Conv2d(C, C, 3, padding=1)).to(dev)
    for p_ in backbone.parameters():  # frozen, like model.backbone in the reference: no gradient, no bucket
        p_.requires_grad = False
    heads = Heads(C, K, J, V, width_for(args.head_params, C, K, V)).to(dev)
    n_train = sum(p_.numel() for p_ in heads.parameters())
    model = nn.parallel.DistributedDataParallel(heads, device_ids=[local]) if world > 1 else heads
    opt = torch.optim.AdamW(heads.parameters(), lr=1e-4)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)  # each rank owns its own frames
    B = args.batch
    to_map = torch.tensor([HM / Wi, HM / Hi], device=dev)
    st = torch.cuda.current_stream(dev)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        return e

    def step(sync=True):
        images = torch.randn((B * V, 3, 128, 128), generator=g, device=dev)
        gt_q = (torch.rand((B, J), generator=g, device=dev) * 2 - 1) * 120.0          # degrees
        e = [ev()]
        with torch.no_grad():                                                         # frozen backbone + GT on the device
            feats = backbone(images)
            gt_uv = mvgeo.project_points(mvgeo.forward_kinematics(chain, gt_q, Rv), rig)   # (B,V,K,2) image px
            gt_kp = (gt_uv * to_map).reshape(B * V, K, 2)                                   # map px
        e.append(ev())
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
            maps, q = model(feats)
        e.append(ev())
        if args.kp_metric == "fused":      # loss + hard decode of the same maps, read once
            l_kpt, dec = mvgeo.decode_and_mse(maps, gt_kp, sigma=5.0, weight=1e4)
        else:
            l_kpt = mvgeo.heatmap_mse_loss(maps, gt_kp, sigma=5.0, weight=1e4)
            dec = mvgeo.decode_heatmaps(maps.detach(), soft=None) if args.kp_metric == "separate" else None
        px_err = (dec.kp_hard - gt_kp).norm(dim=-1).mean() if dec is not None else torch.zeros((), device=dev)
        l_fk, _, _, _ = mvgeo.fk_reproj_loss(chain, q.float(), rig, gt_uv, Rv, lam=1e-4)
        l_ang = nn.functional.smooth_l1_loss(q.float(), gt_q)
        loss = l_kpt + l_fk + l_ang
        e.append(ev())
        opt.zero_grad(set_to_none=True)
        ctx = model.no_sync() if (world > 1 and not sync) else contextlib.nullcontext()
        with ctx:
            loss.backward()                                                           # DDP all-reduce here
        e.append(ev())
        opt.step()
        e.append(ev())
        return e, (loss.detach(), l_kpt.detach(), l_fk.detach(), px_err)

    def timed(n, sync=True):
        rows, hist = [], []
        for _ in range(n):
            e, h = step(sync)
            rows.append(e)
            hist.append(h)
        torch.cuda.synchronize(dev)
        ms = [[a.elapsed_time(b) for a, b in zip(e[:-1], e[1:])] + [e[0].elapsed_time(e[-1])] for e in rows]
        return [statistics.median(c) for c in zip(*ms)], hist

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    phases, hist = timed(args.steps, sync=True)
    nosync, _ = timed(max(5, args.steps // 2), sync=False) if world > 1 else (phases, None)

    # the loss kernels alone: forward and backward of (heat-map MSE + FK reprojection) on detached leaves
    feats = torch.randn((B * V, C, 32, 32), generator=g, device=dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
        maps0, q0 = heads(feats)
    gt_q = (torch.rand((B, J), generator=g, device=dev) * 2 - 1) * 120.0
    gt_uv = mvgeo.project_points(mvgeo.forward_kinematics(chain, gt_q, Rv), rig)
    gt_kp = (gt_uv * to_map).reshape(B * V, K, 2)
    lk_f, lk_b = [], []
    for _ in range(10):
        m_, q_ = maps0.detach().requires_grad_(True), q0.float().detach().requires_grad_(True)
        a = ev()
        l = mvgeo.heatmap_mse_loss(m_, gt_kp, sigma=5.0, weight=1e4) + mvgeo.fk_reproj_loss(chain, q_, rig, gt_uv, Rv, lam=1e-4)[0]
        b = ev()
        l.backward()
        c = ev()
        torch.cuda.synchronize(dev)
        lk_f.append(a.elapsed_time(b))
        lk_b.append(b.elapsed_time(c))
    # the same forward with the per-step key-point metric: one read (fused) against two (MSE, then decode)
    fwd_variants = {}
    for name in ("mse_only", "mse_then_decode", "fused_decode_mse"):
        ts = []
        for _ in range(10):
            a = ev()
            if name == "fused_decode_mse":
                mvgeo.decode_and_mse(maps0, gt_kp, sigma=5.0, weight=1e4)
            else:
                mvgeo.heatmap_mse_loss(maps0, gt_kp, sigma=5.0, weight=1e4)
                if name == "mse_then_decode":
                    mvgeo.decode_heatmaps(maps0, soft=None)
            b = ev()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        fwd_variants[name] = statistics.median(ts)
    ar_ms = None
    if world > 1:  # a stand-alone all-reduce of the trainable gradient bytes
        flat = torch.zeros(n_train, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        ts = []
        for _ in range(10):
            a = ev()
            dist.all_reduce(flat)
            b = ev()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ar_ms = statistics.median(ts)
    vals = torch.tensor(phases + nosync + [statistics.median(lk_f), statistics.median(lk_b), ar_ms or 0.0], device=dev,
                        dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    vals = [float(x) for x in vals]
    p, ns = vals[:6], vals[6:12]
    first, last = [float(x) for x in hist[0]], [float(x) for x in hist[-1]]
    if rank == 0:
        map_bytes = B * V * K * HM * HM * (2 if args.bf16 else 4)
        print(json.dumps({
            "config": "C4: Fr5 DDP training step (heat-map MSE + DH reprojection loss fwd/bwd, NCCL grad all-reduce)",
            "n_gpus": world, "frames_per_gpu": B, "views": V, "steps": args.steps, "heatmap_dtype": "bf16" if args.bf16 else "f32",
            "trainable_params": n_train, "grad_bytes": 4 * n_train,
            "ms_per_step": p[5], "frames_per_s": B * world / (p[5] * 1e-3),
            "ms": {"backbone_and_gt": p[0], "heads_fwd": p[1], "loss_fwd": p[2], "backward_incl_allreduce": p[3], "optimizer": p[4]},
            "ms_no_grad_sync": {"backward": ns[3], "step": ns[5]},
            "allreduce": {"exposed_ms": max(0.0, p[3] - ns[3]), "share_of_step": max(0.0, p[3] - ns[3]) / p[5],
                          "standalone_ms": vals[14] or None,
                          "standalone_busbw_gbs": (4 * n_train * 2 * (world - 1) / world / (vals[14] * 1e-3) / 1e9) if vals[14] else None},
            "loss_kernels": {"fwd_ms": vals[12], "bwd_ms": vals[13], "share_of_step": (vals[12] + vals[13]) / p[5],
                             "pred_map_bytes": map_bytes,
                             "fwd_gbs": map_bytes / (vals[12] * 1e-3) / 1e9, "bwd_gbs": 2 * map_bytes / (vals[13] * 1e-3) / 1e9,
                             "kpt_fwd_ms_rank0": fwd_variants,
                             "kpt_fwd_gbs_one_read_rank0": map_bytes / (fwd_variants["fused_decode_mse"] * 1e-3) / 1e9},
            "kp_metric": args.kp_metric, "kp_px_error_first_last": [first[3], last[3]],
            "loss_first": first, "loss_last": last}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
