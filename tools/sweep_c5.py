"""Config 5: HBM-roofline sweep of the decoder over views V and key-point count K at 480x640 bf16
(decode only: n_maps = B*V*K maps, any K). Prints one JSON line per (V, K)."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvgeo

dev = "cuda:0"
lib = mvgeo._lib.load()
H, W = 480, 640
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
g = torch.Generator(device=dev); g.manual_seed(7)
for V in (2, 4, 8):
    for K in (7, 8, 16, 32):
        B = max(8, (8 * 1024 ** 3) // (V * K * H * W * 2))        # ~8 GB resident pool per (V, K): >> L2
        n_maps = B * V * K
        kp = torch.rand((n_maps, 2), generator=g, device=dev) * torch.tensor([W - 1.0, H - 1.0], device=dev)
        maps = mvgeo.encode_gaussian(kp, (H, W), 3.0, torch.bfloat16)
        idx = torch.empty((n_maps,), dtype=torch.int32, device=dev)
        peak_o, score = torch.empty((n_maps,), device=dev), torch.empty((n_maps,), device=dev)
        kph, kps = torch.empty((n_maps, 2), device=dev), torch.empty((n_maps, 2), device=dev)
        def run():
            rc = lib.mvgeo_decode(maps.data_ptr(), 1, n_maps, H, W, 3.0, 2.5, 1, 100.0, 0, 0, 1, 1, 0, idx.data_ptr(),
                                  peak_o.data_ptr(), score.data_ptr(), kph.data_ptr(), kps.data_ptr(), st)
            assert rc == 0
        for _ in range(3): run()
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts); nbytes = maps.numel() * 2
        ok = bool(((kps - kp * torch.tensor([3.0, 2.5], device=dev)).abs().max() < 3.0))   # decoded == encoded centres
        print(json.dumps({"V": V, "K": K, "frames": B, "GB": round(nbytes / 1e9, 2), "ms": round(ms, 3),
                          "frames_per_s": round(B / ms * 1e3), "GBps": round(nbytes / ms / 1e6), "frac_of_measured_peak": round(nbytes / ms / 1e6 / peak, 3),
                          "frac_of_8TBps": round(nbytes / ms / 1e6 / 8000, 3), "round_trip_ok": ok}), flush=True)
        del maps
