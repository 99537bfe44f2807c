#!/usr/bin/env python
"""A few small steps through every fused-kernel family -- a crash test (an out-of-range shared-memory or global access
faults the launch), also usable under `compute-sanitizer --tool memcheck` where that tool is available (it is closed on
the build pool of round 1): both kernels, both flavours, 2-D mode, optional critics, a DEM with NaN cells (the shared-memory
tile's clamped addressing), a rover at the edge of the map (clamped global path), a rover batch, the closed loop."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from mppi_b200 import capi
    from mppi_b200.core import Core, make_state
    from util import terrain
    dem, cm, hw = terrain("small")
    dem_t, cm_t = torch.from_numpy(dem).cuda(), torch.from_numpy(cm).cuda()
    holes = dem.copy()
    holes[100:140, 120:150] = np.nan
    holes_t = torch.from_numpy(holes).cuda()
    K, T = 256, 24
    n = np.full(T, 0.6, np.float32)
    done = 0
    for math, variants in (("strict", (capi.VARIANT_MONO, capi.VARIANT_PIPE)), ("fast", (capi.VARIANT_PIPE,))):
        for variant in variants:
            for kw in ({}, dict(cw_slope_path=50.5, cw_roll=400.0, cw_pitch=250.0, cw_effort=1.0, cw_orient=1.0,
                                cw_goal_angle=5.0)):
                core = Core(K, T, math=math, variant=variant, **kw)
                for d in (dem_t, holes_t):
                    core.set_terrain(d, hw, cm_t)
                    for st in (make_state(-5.0, -4.0, (0.6, 0.8, 0.0), goal_x=8.0, goal_y=9.0),
                               make_state(hw - 0.3, -hw + 0.2, (1.0, 0.0, 0.0), goal_x=0.0, goal_y=0.0),
                               make_state(1.0, 1.0, (0.0, 1.0, 0.0), goal_x=1.2, goal_y=1.1)):
                        core.set_nominal(n, n)
                        for proj in (capi.PROJ_3D, capi.PROJ_2D):
                            core.step(st, proj, None, 5, done)
                            done += 1
                st = make_state(-5.0, -4.0, (0.6, 0.8, 0.0), goal_x=-4.0, goal_y=-3.0)
                core.set_terrain(dem_t, hw, cm_t)
                core.run_closed_loop(st, 5, capi.PROJ_3D, 3, 0)
                torch.cuda.synchronize()
                core.close()
    R = 6
    core = Core(K, T, max_rovers=R)
    core.set_terrain_batched(dem_t[None].repeat(R, 1, 1), hw, cm_t[None].repeat(R, 1, 1))
    states = core.pack_states([make_state(-5.0 + r, -4.0, (0.6, 0.8, 0.0), goal_x=8.0, goal_y=9.0) for r in range(R)],
                              core.device)
    core.step_batched(states, R, seed=1, offset=2)
    torch.cuda.synchronize()
    core.close()
    print(f"sanitize_smoke: {done} steps ok")


if __name__ == "__main__":
    main()
