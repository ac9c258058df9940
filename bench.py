#!/usr/bin/env python
"""bench.py — frames/s of decode + triangulate + FK on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # N > 1: one rank per GPU, weak scaling

Workload (config 2 of BASELINE.json): FR3, 4 views, 1024 frames per GPU per step, 8 key-points,
240x320 bf16 belief maps = 5.03 GB per step per GPU (>> 126 MB L2, so no flush is needed
between steps). Synthetic closed-loop data: joint angles -> FK -> projection -> Gaussian blobs
(sigma 3 px) + N(0, 0.01) noise, generated on the device before the timed region.

A step = one pass of the hot path over one batch: belief-map decode (arg-max + global
soft-arg-max) -> DLT triangulation -> FK + reprojection consistency, then (N > 1) the final
NCCL result gather. `value` times K steps with CUDA events on the launching stream, inputs
resident in HBM; `e2e` times the host-buffer C-ABI call (pinned host maps in, host results
out, H2D/D2H inside the timed region). `roofline` is the decode kernel: algorithmic bytes
(V*K*H*W*2 per frame) over its CUDA-event duration inside the same timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

#: BASELINE.json configs. The bench line is C2 (the configuration the metric is quoted on); the others
#: are selectable with --workload for the tables in DESIGN.md (they are parity-test cases otherwise).
WORKLOADS = {
    "c1": dict(robot="fr3", V=3, B=8, H=120, W=160, dtype="f32", name="C1: FR3 3-view, batch 8, 120x160 fp32 belief maps"),
    "c2": dict(robot="fr3", V=4, B=1024, H=240, W=320, dtype="bf16", name="C2: FR3 4-view, batch 1024 frames/GPU, 240x320 bf16 belief maps"),
    "c3": dict(robot="meca500", V=4, B=2048, H=240, W=320, dtype="bf16", name="C3: Meca500 4-camera, 2048-frame resident chunk/GPU of the 65,536-frame job, 240x320 bf16"),
    "c5": dict(robot="fr3", V=8, B=128, H=480, W=640, dtype="bf16", name="C5: FR3 8-view, 128-frame resident chunk/GPU, 480x640 bf16 belief maps"),
}
ROBOT, V, B, H, W, MAP_DTYPE = "fr3", 4, 1024, 240, 320, "bf16"
BETA, MIN_SCORE = 100.0, 0.5
METRIC = "frames/s decode+triangulate+FK"
WORKLOAD = WORKLOADS["c2"]["name"] + ", decode+triangulate+FK"


def select_workload(key: str):
    global ROBOT, V, B, H, W, MAP_DTYPE, WORKLOAD
    w = WORKLOADS[key]
    ROBOT, V, B, H, W, MAP_DTYPE = w["robot"], w["V"], w["B"], w["H"], w["W"], w["dtype"]
    WORKLOAD = w["name"] + ", decode+triangulate+FK"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per decode launch from the committed
    `ncu --set full` capture of this same command (profiles/r01_ncu_decode_raw_selected.csv)."""
    import csv

    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_decode_raw_selected.csv")) as f:
            rows = list(csv.reader(f))
        hdr, units, first = rows[0], rows[1], rows[2]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        return float(first[rd]) * scale[units[rd]] + float(first[wr]) * scale[units[wr]]
    except Exception:
        return None


def _intrinsics():
    import mvgeo

    return [[[v[0], 0.0, v[2]], [0.0, v[1], v[3]], [0.0, 0.0, 1.0]] for v in mvgeo.ZEDX_FHD1200.values()]


def _config(world: int, n_keypoints: int) -> dict:
    """The `config` object of the JSON line (identical for both arms)."""
    esize = 2 if MAP_DTYPE == "bf16" else 4
    nbytes = V * n_keypoints * H * W * esize * B
    return {"workload": WORKLOAD, "robot": ROBOT, "views": V, "keypoints": n_keypoints, "frames_per_gpu_per_step": B,
            "map": [H, W], "map_dtype": MAP_DTYPE, "soft_argmax": f"global beta={BETA}",
            "l2": f"inputs are {nbytes / 1e9:.2f} GB per step per GPU" +
                  (" (>> 126 MB L2): no flush needed" if nbytes > 4e8 else " (< L2): L2 flushed by a 256 MB write between steps"),
            "result_gather": ("results of every batch (X_tri, kp_soft, score, X_fk) kept in a device ring; ONE final nccl "
                              "all_gather_into_tensor per job, inside the timed region") if world > 1 else "none (1 GPU)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port; the reference itself is
    Python and cannot travel to the GPU box), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_pipeline as cp

    ncores = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        per_step = float(os.environ.get("MVGEO_BENCH_REF_SECONDS", 0)) or max(2.0, min(10.0, 90.0 / max(1, args.warmup + args.steps)))
        fps, workers, total, desc = cp.timed_throughput(ROBOT, V, H, W, (1200, 1920), _intrinsics(), frames_per_worker=4,
                                                        min_seconds=per_step, workers=ncores)
        if i >= args.warmup:
            vals.append(fps)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * B / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(1, {"fr3": 8, "fr5": 7, "meca500": 7}[ROBOT]),
        "note": f"CPU path of the same workload; ms_per_step = time for one {B}-frame batch at the measured rate",
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def make_inputs(mv, torch, dev, rank):
    """Closed-loop synthetic belief maps on the device (untimed)."""
    import numpy as np

    chain = mv.Chain.builtin(ROBOT)
    rig = mv.CameraRig.synthetic_ring_for(ROBOT, V)  # aimed at the arm: every key-point is in view
    views = (list(mv.VIEW_EULER_ZYX_DEG[ROBOT]) + [None] * V)[:V]
    Rv = np.stack([np.asarray(mv.view_rotation(ROBOT, v)) for v in views]).astype(np.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    span = 0.9 * 2.8 if ROBOT == "fr3" else 0.9 * 150.0  # radians for FR3, degrees for Fr5 / Meca500
    q = (torch.rand((B, chain.n_joints), generator=g, device=dev) * 2.0 - 1.0) * span
    X = mv.forward_kinematics(chain, q, Rv)
    uv = mv.project_points(X, rig)
    Hi, Wi = rig.image_size
    kp_map = uv * torch.tensor([W / Wi, H / Hi], device=dev)
    tdt = torch.bfloat16 if MAP_DTYPE == "bf16" else torch.float32
    maps = mv.encode_gaussian(kp_map, (H, W), 3.0, tdt)
    for b0 in range(0, B, 32):  # noise in slices: no 10 GB temporary
        sl = maps[b0:b0 + 32]
        sl.add_(torch.randn(sl.shape, generator=g, device=dev, dtype=torch.float32).mul_(0.01).to(tdt))
    P = torch.from_numpy(rig.projection_matrices(Rv.astype(np.float64))).to(dev)
    return chain, rig, Rv, q, maps, P


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mvgeo
    from mvgeo import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = mvgeo._lib.load()

    chain, rig, Rv, q, maps, P = make_inputs(mvgeo, torch, dev, rank)
    K = chain.n_points
    cams = ops.cameras_to_device(rig, dev)
    Rvt = torch.from_numpy(Rv).to(dev)
    # Every batch of the job writes its per-frame results (everything that leaves the GPU) into its own
    # slot of one device ring; the job's ONLY collective is the final gather of that ring.
    spec = {"X_tri": ((B, K, 3), torch.float32), "kp_soft": ((B, V, K, 2), torch.float32),
            "score": ((B, V, K), torch.float32), "X_fk": ((B, V, K, 3), torch.float32)}
    n_slots = min(max(args.steps, 1), 128)
    ring = mvgeo.sharding.ResultRing(spec, n_slots, dev)
    scratch = mvgeo.alloc_outputs(B, V, K, dev)  # outputs that stay on the GPU (idx, peak, residuals, loss ...)
    counter = [0]
    out = dict(scratch)
    flush = torch.zeros(64 * 1024 * 1024, device=dev) if maps.numel() * maps.element_size() < 4e8 else None
    Hi, Wi = rig.image_size
    sx, sy = Wi / W, Hi / H
    n_maps = B * V * K
    DT = mvgeo._lib.BF16 if MAP_DTYPE == "bf16" else mvgeo._lib.F32
    esize = 2 if MAP_DTYPE == "bf16" else 4
    st = torch.cuda.current_stream(dev)
    import ctypes as C

    def step():
        """decode -> (triangulate || FK + reprojection consistency + loss sum) through the C ABI: the same
        two launches mvgeo_pipeline makes, with CUDA events around the decode kernel."""
        j = counter[0] % n_slots
        if world > 1 and counter[0] > 0 and j == 0:
            ring.final_gather()  # ring full (more than 128 batches in the job): flush before re-use
        counter[0] += 1
        out = dict(scratch)
        for name in spec:
            out[name] = ring.slot[j][name]
        if flush is not None:
            flush.add_(1.0)  # inputs smaller than L2: evict them between steps (untimed by the decode events)
        s = st.cuda_stream
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = lib.mvgeo_decode(maps.data_ptr(), DT, n_maps, H, W, sx, sy, mvgeo._lib.SOFT_GLOBAL, BETA, 0, 0,
                              1, 1, 0, out["idx"].data_ptr(), out["peak"].data_ptr(), out["score"].data_ptr(),
                              out["kp_hard"].data_ptr(), out["kp_soft"].data_ptr(), s)
        e1.record(st)
        rc |= lib.mvgeo_geometry(out["kp_soft"].data_ptr(), out["score"].data_ptr(), P.data_ptr(), C.byref(chain.struct),
                                 q.data_ptr(), B, Rvt.data_ptr(), cams.data_ptr(), V, K, MIN_SCORE, 0, 1.0,
                                 out["X_tri"].data_ptr(), out["tri_resid"].data_ptr(), out["tri_views"].data_ptr(),
                                 out["X_fk"].data_ptr(), out["uv_fk"].data_ptr(), out["frame_loss"].data_ptr(),
                                 out["loss"].data_ptr(), out["ticket"].data_ptr(), s)
        assert rc == 0, rc
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record(st)
        return e0, e1, e2

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def drain():
        """The path's only communication: one final gather of the job's results, < 1 KB per frame."""
        if world > 1:
            used = counter[0] % n_slots or min(counter[0], n_slots)
            ring.final_gather(used)
        counter[0] = 0

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    if world > 1:  # warm the collective up at the message size the timed region will use
        for _ in range(2):
            ring.final_gather(min(args.steps, n_slots))
    # Timed region; measured again (once) if the clocks were throttled by hw_slowdown / thermal events.
    BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    remeasured = False
    for attempt in range(2):
        sampler = ClockSampler(local)  # NVML init takes milliseconds: do it BEFORE the barrier so that every
        sampler.start()                # rank enters the timed region together
        fence()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record(st)
        t_cpu0 = time.perf_counter()
        dec_events = [step() for _ in range(args.steps)]
        cpu_issue_ms = 1e3 * (time.perf_counter() - t_cpu0) / args.steps  # host time to enqueue one step
        t_gather = torch.cuda.Event(enable_timing=True)
        t_gather.record(st)
        drain()  # the final result gather is inside the timed region
        t_end.record(st)
        fence()
        clocks = sampler.stop()
        throttled = torch.tensor([1.0 if BAD & set(clocks["reasons"]) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(throttled, op=dist.ReduceOp.MAX)  # every rank takes the same decision
        if float(throttled) == 0.0 or attempt == 1:
            break
        remeasured = True
    clocks["remeasured"] = remeasured
    elapsed_ms = t_start.elapsed_time(t_end)
    dec_ms = [a.elapsed_time(b) for a, b, _ in dec_events]
    if flush is not None:  # small workload: the L2 flush between steps is not part of the path
        elapsed_ms = sum(a.elapsed_time(c) for a, _, c in dec_events) + t_gather.elapsed_time(t_end)
    gather_ms = t_gather.elapsed_time(t_end)
    out = dict(scratch)
    for name in spec:
        out[name] = ring.slot[(args.steps - 1) % n_slots][name]
    loss = float(out["loss"])
    assert np.isfinite(loss)
    frac_all_views = float((out["tri_views"] == V).float().mean())

    # ---------------- e2e: host buffers through the C-ABI context (H2D + kernels + D2H timed)
    e2e_steps, e2e_s, h2d, d2h = 0, float("nan"), 0, 0
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))
        hp = mvgeo.HostPipeline(chain, rig, Rv, dtype=maps.dtype, H=H, W=W, image_size=rig.image_size, soft="global",
                                beta=BETA, min_score=MIN_SCORE, chunk_frames=64, device=local)
        maps_h = torch.empty(maps.shape, dtype=maps.dtype).pin_memory()
        maps_h.copy_(maps)
        q_h = q.cpu().pin_memory()
        out_h = mvgeo.alloc_outputs(B, V, K, None, pin=True)
        hp.run(maps_h, q_h, out_h)  # warm-up
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hp.run(maps_h, q_h, out_h)  # synchronous on return: results are in host memory
        e2e_s = time.perf_counter() - t0
        assert np.isfinite(float(out_h["loss"]))
        assert torch.equal(out_h["idx"], out["idx"].cpu()), "host pipeline and device pipeline disagree"
        hp.close()
        h2d = maps_h.numel() * maps_h.element_size() + q_h.numel() * 4 + V * (12 + 9 + 24) * 4
        d2h = sum(t.numel() * t.element_size() for n, t in out_h.items() if isinstance(t, torch.Tensor) and n != "loss")

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s, statistics.mean(dec_ms), gather_ms, cpu_issue_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s, dec_mean, gather_ms, cpu_issue_ms = (float(x) for x in t)
    else:
        dec_mean = statistics.mean(dec_ms)

    if rank == 0:
        peak, peak_src = _peaks()
        frame_bytes = V * K * H * W * esize
        achieved = frame_bytes * B / (dec_mean * 1e-3) / 1e9
        value = B * world * args.steps / (elapsed_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(world, K),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _profiled_traffic() if args.workload == "c2" else None, "algorithmic_bytes": frame_bytes * B,
                         "kernel": f"decode_tma_kernel<{MAP_DTYPE}, global, persistent>", "peak_source": peak_src,
                         "decode_ms": dec_mean, "decode_share_of_step": dec_mean * args.steps / elapsed_ms,
                         "frac_of_nominal_8TBps": achieved / 8000.0, "bytes_per_frame": frame_bytes},
            "e2e": {"value": (B * world * e2e_steps / e2e_s) if e2e_steps else None, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": "mvgeo_pipeline_host (pinned host buffers)"},
            "gpu_launches": 2 * args.steps,
            "breakdown": {"final_gather_ms": gather_ms, "host_enqueue_ms_per_step": cpu_issue_ms,
                          "result_bytes_per_step_per_gpu": 4 * ring.record},
            "clocks": clocks,
            "check": {"loss_px2": loss, "rms_reproj_px": loss ** 0.5, "frames_with_all_views": frac_all_views},
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import cpu_pipeline as cp

            fps, workers, total, desc = cp.timed_throughput(ROBOT, V, H, W, (1200, 1920), _intrinsics(), frames_per_worker=4,
                                                            min_seconds=10.0)
            fps1, _, _, _ = cp.timed_throughput(ROBOT, V, H, W, (1200, 1920), _intrinsics(), frames_per_worker=4,
                                                min_seconds=4.0, workers=1)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": workers, "kind": "port", "sample": desc,
                                    "single_process_value": fps1,
                                    "note": "value = one process per host core (the reference's DataLoader-worker analogue); "
                                            "single_process_value = one Python process, exactly how the reference loops"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (bench line: c2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
