"""Synthetic lunar terrain of the shapes BASELINE.json names (the reference's data blobs are absent:
/root/reference/.MISSING_LARGE_BLOBS).  Recipes follow the reference's own generators:

  DEM      crater field of Surface.create_surface (MPPI_isaac.py:317-320); the nine craters of the 150 m
           map are the commented list in MPPI_OO_current.py:730-740, larger maps draw craters from
           default_rng(57) (the driver's terrain seed, visual_terrain_stack_full_terrain.py:11).
  costmap  750 random rocks from RandomState(99) (MPPI_OO_current.py:721-725), inflated discs
           r + r_robot + 0.2 (MPPI_OO_current.py:293), cv2.distanceTransform(DIST_L2, 5), min-max
           normalise, (1 - d)^10 (create_costmap.py:14-28).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NINE_CRATERS = [
    ((-2.7, -19.0), 3.4, 12.23), ((-0.57, -0.05), 4.39, 11.52), ((-48.56, 12.78), 3.6, 12.4),
    ((-27.89, 38.56), 4.0, 12.7), ((-50.12, 19.34), 3.7, 13.0), ((20.45, -48.78), 4.4, 12.9),
    ((-20.67, -40.12), 4.2, 12.9), ((42.78, 21.56), 4.5, 12.7), ((-36.12, -33.34), 3.9, 13.0),
]


def random_craters(half_width: float, seed: int = 57, density_per_m2: float = 1.0 / 625.0):
    rng = np.random.default_rng(seed)
    n = max(1, int((2 * half_width) ** 2 * density_per_m2))
    return [((float(rng.uniform(-half_width, half_width)), float(rng.uniform(-half_width, half_width))),
             float(rng.uniform(0.5, 4.5)), float(rng.uniform(2.0, 13.0))) for _ in range(n)]


def crater_dem(grid_size: int, half_width: float, bumps=None, seed: int = 57, device="cpu"):
    """float32 [grid_size, grid_size] crater field.  Uses torch so that large maps can be built on the GPU."""
    import torch
    if bumps is None:
        bumps = NINE_CRATERS if abs(half_width - 75.0) < 1e-9 else random_craters(half_width, seed)
    x = torch.linspace(-half_width, half_width, grid_size, dtype=torch.float64, device=device)
    Z = torch.zeros((grid_size, grid_size), dtype=torch.float64, device=device)
    res = 2 * half_width / (grid_size - 1)
    for (cx, cy), h, w in bumps:
        # craters have compact numerical support: restrict the update to +-6 w
        r = 6.0 * w
        i0, i1 = max(0, int((cx - r + half_width) / res)), min(grid_size, int((cx + r + half_width) / res) + 2)
        j0, j1 = max(0, int((cy - r + half_width) / res)), min(grid_size, int((cy + r + half_width) / res) + 2)
        if i0 >= i1 or j0 >= j1:
            continue
        dx2 = (x[i0:i1] - cx) ** 2
        dy2 = (x[j0:j1] - cy) ** 2
        r2 = dy2[:, None] + dx2[None, :]
        Z[j0:j1, i0:i1] += (h - 0.5) * torch.exp(-r2 / (2 * w ** 2)) - (h + 0.5) * torch.exp(-r2 / (2 * (w / 2) ** 2))
    return Z.to(torch.float32)


def rock_free_mask(costmap_size: int, half_width: float, n_rocks: int = 750, seed: int = 99,
                   r_robot: float = 0.3) -> np.ndarray:
    """uint8 [costmap_size, costmap_size]: 255 = free, 0 = inside a rock disc inflated by the robot radius + 0.2 m
    (rocks: RandomState(99), centre U(-2/3 hw, 2/3 hw)^2, r in U(0, 0.4); MPPI_OO_current.py:723-725, :293)."""
    rng = np.random.RandomState(seed)
    span = half_width * 2.0 / 3.0
    xc = np.linspace(-half_width, half_width, costmap_size)
    res = 2 * half_width / (costmap_size - 1)
    free = np.full((costmap_size, costmap_size), 255, dtype=np.uint8)
    for _ in range(n_rocks):
        ox, oy, r = rng.uniform(-span, span), rng.uniform(-span, span), rng.uniform(0.0, 0.4)
        R = r + r_robot + 0.2
        i0, i1 = max(0, int((ox - R + half_width) / res)), min(costmap_size, int((ox + R + half_width) / res) + 2)
        j0, j1 = max(0, int((oy - R + half_width) / res)), min(costmap_size, int((oy + R + half_width) / res) + 2)
        sub = (xc[None, i0:i1] - ox) ** 2 + (xc[j0:j1, None] - oy) ** 2 <= R * R
        free[j0:j1, i0:i1][sub] = 0
    return free


def costmap_from_free_mask(free: np.ndarray, power: float = 10.0) -> np.ndarray:
    """The reference's offline recipe (create_costmap.py:14-28): L2 distance transform (5x5 mask) of the free space,
    min-max normalised to [0, 1], cost = (1 - d)^power."""
    import cv2
    dist = cv2.distanceTransform(free, cv2.DIST_L2, 5)
    dist = cv2.normalize(dist, None, 0, 1.0, cv2.NORM_MINMAX)
    return ((1.0 - dist) ** power).astype(np.float32)


def rock_costmap(costmap_size: int, half_width: float, n_rocks: int = 750, seed: int = 99, r_robot: float = 0.3,
                 power: float = 10.0) -> np.ndarray:
    """float32 [costmap_size, costmap_size] obstacle costmap in [0, 1]."""
    return costmap_from_free_mask(rock_free_mask(costmap_size, half_width, n_rocks, seed, r_robot), power)


@dataclass
class Workload:
    name: str
    K: int
    T: int
    grid_size: int
    half_width: float
    costmap_size: int
    n_rovers: int = 1
    start: tuple = (-60.57, -60.23)
    goal: tuple = (65.80, 65.40)

    @property
    def scale(self):
        return self.half_width / 75.0


# BASELINE.json configs.  Start/goal are those of MPPI_OO_current.py:831-835 scaled to the map.
WORKLOADS = {
    "C1": Workload("C1 cpu-reference K=1024 T=50 DEM1500 costmap750", 1024, 50, 1500, 75.0, 750),
    "C2": Workload("C2 K=4096 T=100 DEM1500 costmap750", 4096, 100, 1500, 75.0, 750),
    "C3": Workload("C3 K=262144 T=100 DEM2048 costmap1024", 262144, 100, 2048, 102.4, 1024),
    "C4": Workload("C4 4096 rovers x K=1024 T=64 DEM512 costmap256", 1024, 64, 512, 25.6, 256, n_rovers=4096),
    "C5": Workload("C5 K=65536 T=200 DEM8192 costmap1024", 65536, 200, 8192, 102.4, 1024),
}


def workload_start_goal(w: Workload):
    s = w.scale
    return (w.start[0] * s, w.start[1] * s), (w.goal[0] * s, w.goal[1] * s)
