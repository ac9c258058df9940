"""CUDA-graph capture of the fused pipeline for fixed shapes: the three launches (decode,
triangulate, FK + consistency) replay as one graph launch, which is what a latency-bound caller
(the reference's per-frame real-time loop, DIP_REAL.py:98-127, or small batches such as BASELINE
config 1) needs — the kernels themselves take a few microseconds there."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .robots import Chain


class GraphedPipeline:
    """Capture `ops.pipeline` once for static input buffers; `run()` copies nothing and allocates
    nothing: write new data into `.maps` / `.q` (device tensors) and replay.

        gp = GraphedPipeline(chain, rig, R_view, batch=8, dtype=torch.float32, H=120, W=160, image_size=(1200, 1920))
        gp.maps.copy_(new_maps); gp.q.copy_(new_q); out = gp.run()     # dict of device tensors (static buffers)
    """

    def __init__(self, chain: Chain, rig, R_view, *, batch: int, dtype, H: int, W: int, image_size=None,
                 device="cuda", **pipeline_kwargs):
        import numpy as np

        dev = torch.device(device)
        V, K = rig.n_views, chain.n_points
        self.maps = torch.zeros((batch, V, K, H, W), dtype=dtype, device=dev)
        self.q = torch.zeros((batch, chain.n_joints), dtype=torch.float32, device=dev)
        Rv = None if R_view is None else np.asarray(R_view, dtype=np.float64)
        self._P = torch.from_numpy(rig.projection_matrices(Rv)).to(dev)
        self._cams = ops.cameras_to_device(rig, dev)
        self._Rv = None if R_view is None else torch.as_tensor(np.asarray(R_view, dtype=np.float32)).to(dev)
        self.out = ops.alloc_outputs(batch, V, K, dev)
        self._args = dict(image_size=image_size, out=self.out, **pipeline_kwargs)
        self._chain = chain
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture (function attributes, lazy module load)
            self._launch()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._launch()

    def _launch(self):
        ops.pipeline(self.maps, self._P, self._chain, self.q, self._cams, self._Rv, **self._args)

    def run(self) -> dict:
        self.graph.replay()
        return self.out
