#!/usr/bin/env python
"""Times the obstacle-costmap rebuild (f1): GPU builder vs the reference recipe (NumPy mask per rock + cv2) on the host.

  python tools/costmap_bench.py [--size 875] [--half-width 87.5] [--rocks 750]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=875)            # Isaac run: grid 7000 / 8 (MPPI_isaac.py:271)
    ap.add_argument("--half-width", type=float, default=87.5)
    ap.add_argument("--rocks", type=int, default=750)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import cv2
    import torch
    from mppi_b200 import build_obstacle_costmap
    rng = np.random.default_rng(5)
    span = 0.7 * a.half_width
    obst = [(float(rng.uniform(-span, span)), float(rng.uniform(-span, span)), float(rng.uniform(0.1, 0.8)))
            for _ in range(a.rocks)]
    n, hw = a.size, a.half_width
    out = torch.empty((n, n), dtype=torch.float32, device="cuda")
    for _ in range(3):
        build_obstacle_costmap(obst, (0.0, 0.0), n, hw, 0.3, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        build_obstacle_costmap(obst, (0.0, 0.0), n, hw, 0.3, out=out)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        build_obstacle_costmap(obst, (0.0, 0.0), n, hw, 0.3, out=out)
    e1.record()
    torch.cuda.synchronize()
    # reference recipe on the host (MPPI_isaac.py:361-378) + the H2D copy the driver does (:563)
    t0 = time.perf_counter()
    xc = np.linspace(-hw, hw, n)
    X, Y = np.meshgrid(xc, xc)
    obs = 255 * np.ones((n, n), dtype=np.uint8)
    for xg, yg, r in obst:
        R = r / 2 + 0.3 + 0.1
        obs[(X - yg) ** 2 + (Y - xg) ** 2 <= R ** 2] = 0
    t1 = time.perf_counter()
    d = cv2.distanceTransform(obs, cv2.DIST_L2, 5)
    dn = cv2.normalize(d, None, 0, 1.0, cv2.NORM_MINMAX)
    c = ((1 - dn) ** 20).astype(np.float32)
    t2 = time.perf_counter()
    out.copy_(torch.from_numpy(c))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    got = out.cpu().numpy()
    build_obstacle_costmap(obst, (0.0, 0.0), n, hw, 0.3, out=out)
    torch.cuda.synchronize()
    print(json.dumps({"costmap": f"{n}x{n}", "rocks": a.rocks,
                      "gpu_ms_host_wall_median": float(np.median(ts) * 1e3),
                      "gpu_ms_device": float(e0.elapsed_time(e1) / a.reps),
                      "reference_ms": {"numpy_masks": (t1 - t0) * 1e3, "cv2_dt_normalize_pow": (t2 - t1) * 1e3,
                                       "h2d": (t3 - t2) * 1e3, "total": (t3 - t0) * 1e3},
                      "max_abs_diff_vs_reference": float(np.max(np.abs(out.cpu().numpy() - got)))}))


if __name__ == "__main__":
    main()
