#!/usr/bin/env python
"""Smallest program that launches the fused kernel of one BASELINE workload a few times -- the command profiled with
ncu (profiles/r2_ncu_*.md).  No timing here: numbers taken under a profiler are never bench numbers.

  python tools/prof_run.py --workload C2|C3|C4|C5|C5many [--steps 6] [--math strict] [--rovers 512]

C5many: BASELINE config 5's DEM (8192^2 fp32, 268 MB) with 64 controllers started on an 8 x 8 grid of poses spread over
the map (K = 1024 each = 65536 samples, T = 200, one shared DEM and costmap, mppi_step_batched): the touched windows add
up to ~530 MB, so the terrain gathers really leave L2 (SURVEY.md 8d, C5 caveat)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def many_start_states(w, n_side=8):
    from mppi_b200.core import make_state
    span = w.half_width * 0.75
    xs = np.linspace(-span, span, n_side)
    states = []
    for j, y in enumerate(xs):
        for i, x in enumerate(xs):
            a = 2 * np.pi * ((i * n_side + j) % 16) / 16.0
            states.append(make_state(float(x), float(y), (float(np.cos(a)), float(np.sin(a)), 0.0),
                                     goal_x=float(-x), goal_y=float(-y)))
    return states


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--math", default="strict")
    ap.add_argument("--rovers", type=int, default=512)
    ap.add_argument("--flush", action="store_true", help="flush L2 (256 MiB fill) before every step")
    ap.add_argument("--K", type=int, default=0, help="override the workload's samples (single-controller workloads)")
    a = ap.parse_args()
    import torch
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state
    dev = torch.device("cuda", 0)
    name = "C5" if a.workload == "C5many" else a.workload
    w = syn.WORKLOADS[name]
    if a.K:
        import dataclasses
        w = dataclasses.replace(w, K=a.K)
    ext = dict(cw_slope_path=50.5, cw_roll=400.0, cw_pitch=250.0) if name == "C5" else {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if a.flush else None
    if a.workload in ("C4", "C5many"):
        if a.workload == "C4":
            R, K = a.rovers, w.K
            pool = 16
            dem_pool = torch.stack([syn.crater_dem(w.grid_size, w.half_width, seed=57 + i, device=dev) for i in range(pool)])
            cm_pool = torch.stack([torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width, n_rocks=90, seed=99 + i))
                                   for i in range(pool)]).to(dev)
            idx = torch.arange(R, device=dev) % pool
            dems, cms = dem_pool[idx].contiguous(), cm_pool[idx].contiguous()
            rng = np.random.default_rng(7)
            half = w.half_width / 2
            states = [make_state(float(rng.uniform(-half, half)), float(rng.uniform(-half, half)),
                                 (float(np.cos(t)), float(np.sin(t)), 0.0), goal_x=float(rng.uniform(-half, half)),
                                 goal_y=float(rng.uniform(-half, half))) for t in rng.uniform(0, 2 * np.pi, R)]
        else:
            states = many_start_states(w)
            R, K = len(states), 1024
            dem = syn.crater_dem(w.grid_size, w.half_width, device=dev).contiguous()
            cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
            dems, cms = dem.unsqueeze(0).expand(R, -1, -1), cm.unsqueeze(0).expand(R, -1, -1)     # ONE shared map
        core = Core(K, w.T, math=a.math, max_rovers=R, **ext)
        core.set_terrain_batched_shared(dems, w.half_width, cms) if a.workload == "C5many" else \
            core.set_terrain_batched(dems, w.half_width, cms)
        sd = Core.pack_states(states, dev)
        for i in range(a.steps):
            if flush is not None:
                flush.fill_(i)
            core.step_batched(sd, R, capi.PROJ_3D, 42, i)
    else:
        dem = syn.crater_dem(w.grid_size, w.half_width, device=dev).contiguous()
        cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
        start, goal = syn.workload_start_goal(w)
        st = make_state(start[0], start[1], goal_x=goal[0], goal_y=goal[1])
        core = Core(w.K, w.T, math=a.math, **ext)
        core.set_terrain(dem, w.half_width, cm)
        for i in range(a.steps):
            if flush is not None:
                flush.fill_(i)
            core.step(st, capi.PROJ_3D, None, 42, i)
    torch.cuda.synchronize()
    print("prof_run ok", a.workload, core.read_stats(0))


if __name__ == "__main__":
    main()
