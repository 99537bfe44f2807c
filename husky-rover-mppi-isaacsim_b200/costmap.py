"""GPU obstacle-costmap builder: host wrapper over mppi_build_costmap.

Replaces Surface.create_obstacles_costmap (thesis_master/warp_implementation/MPPI_isaac.py:361-378), which the Isaac
driver re-runs at every terrain-block change and then uploads with `costmap_wp.assign(...)`
(visual_terrain_stack_full_terrain.py:561-563).  The result lands directly in a device tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import capi


def build_obstacle_costmap(obstacles: Sequence, origin, costmap_size: int, half_width: float, r_robot: float,
                           out: Optional[torch.Tensor] = None, device: int = 0, power: float = 20.0,
                           radius_scale: float = 0.5, inflate: float = 0.1, want_intermediates: bool = False,
                           stream=None):
    """obstacles: iterable of (x_global, y_global, r_obs).  Returns the device costmap [cms, cms] float32
    (and, with want_intermediates, the distance map and the uint8 mask)."""
    if not torch.cuda.is_available():
        raise capi.MppiError("the costmap builder needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", device)
    obs = np.ascontiguousarray(np.asarray(list(obstacles), dtype=np.float64).reshape(-1, 3))
    n = int(costmap_size)
    if out is None:
        out = torch.empty((n, n), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.dtype == torch.float32 and out.numel() == n * n and out.is_contiguous()
    dist = torch.empty((n, n), dtype=torch.float32, device=dev) if want_intermediates else None
    mask = torch.empty((n, n), dtype=torch.uint8, device=dev) if want_intermediates else None
    s = (stream if stream is not None else torch.cuda.current_stream(dev)).cuda_stream
    capi.check(capi.lib().mppi_build_costmap(device, obs.ctypes.data_as(C.c_void_p), obs.shape[0], float(origin[0]),
                                             float(origin[1]), n, float(half_width), float(r_robot),
                                             float(radius_scale), float(inflate), float(power), out.data_ptr(),
                                             dist.data_ptr() if dist is not None else None,
                                             mask.data_ptr() if mask is not None else None, s), "mppi_build_costmap")
    return (out, dist, mask) if want_intermediates else out
