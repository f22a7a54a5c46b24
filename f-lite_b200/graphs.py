"""CUDA-graph replay of the DiT forward for launch-bound shapes (small models / low resolution).

One denoise step issues ~10 launches per block through ctypes; at the 10B / 1024^2 workload the GPU is the bottleneck
(375 launches of ~280 us each) but at small shapes the host is (config C1: 42 launches of a few microseconds each,
~25 us of Python + ctypes per launch).  ``GraphedForward`` captures ``dit(cat([latents, latents]), context, mask, t)``
once per (shapes, context) into a ``torch.cuda.CUDAGraph`` and replays it every step: the latents are read from the
caller's tensor *in place* (``flite_cfg_euler`` updates that same tensor between replays), the timestep is copied into a
static device tensor, the velocity comes back in a static output buffer.

All libflite_b200 entry points are stream-ordered, allocate nothing and never synchronise, so they are capturable as
they are; the TMA descriptors are by-value kernel parameters and are frozen into the graph together with the workspace
addresses.  Not available with sequence parallelism (the peer-memory handshake bakes a host-incremented epoch into its
kernel parameters).
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import FliteError


class GraphedForward:
    def __init__(self, dit_model, latents: torch.Tensor, context: torch.Tensor, mask: Optional[torch.Tensor],
                 t_example: torch.Tensor, duplicate_latents: bool = True):
        if getattr(dit_model, "sp_group", None) is not None:
            raise FliteError("CUDA-graph replay is not available with sequence parallelism")
        self.model = dit_model
        self.latents = latents                       # read in place at every replay
        self.context, self.mask = context, mask      # kept alive: the graph holds their addresses
        self.t = t_example.clone()
        self.dup = duplicate_latents
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                # warm-up outside the capture: attributes set, workspaces allocated,
            for _ in range(2):                       # hoisted context K/V computed and cached
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.out = self._forward()
        # the graph has the workspace addresses baked in: keep those allocations alive even if the model later grows
        # (and thereby replaces) a workspace for a larger shape
        self._keep = (dict(dit_model._ws), dict(dit_model._rope_cache), dit_model._ctx_cache)

    def _forward(self):
        x = torch.cat([self.latents] * 2) if self.dup else self.latents
        return self.model(x, self.context, self.mask, self.t)

    def __call__(self, t_tensor: torch.Tensor) -> torch.Tensor:
        self.t.copy_(t_tensor, non_blocking=True)
        self.graph.replay()
        return self.out
