#!/usr/bin/env python
"""Times the reference's OWN CPU projection path -- `generate_trajectory_25D` of
thesis_master/python_mppi_projection/displacement_on_surface.py:317-369, executed as written (function block of the
script with `np.int(` -> `int(`, no plotting) -- on this machine's CPU: BASELINE.md's baseline "B0".

  python tools/time_reference_cpu_script.py [--samples 32] [--T 50] > profiles/rNN_cpu_reference_script.json

Build container only (needs /root/reference).  The script rolls out ONE trajectory at a time on one core; a
K = 1024, T = 50 iteration is K calls.  This is the reference's rollout alone (no wheels, critics or update), so it is
a floor for the reference's CPU cost of an MPPI iteration.  bench.py's `cpu_baseline` is the C port of the full
iteration, which is what can travel to the GPU box.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--T", type=int, default=50)
    a = ap.parse_args()
    from make_golden import load_reference_functions
    ns = load_reference_functions()
    gs, hw = 1500, 75.0
    bumps = [((-2.7, -19.0), 3.4, 12.23), ((-0.57, -0.05), 4.39, 11.52), ((-48.56, 12.78), 3.6, 12.4)]
    X, Y, Z = ns["create_surface"](gs, hw, bumps)
    res = 2 * hw / gs
    rng = np.random.default_rng(0)
    times = []
    for k in range(a.samples + 2):
        v = rng.uniform(0.2, 2.0, a.T)
        w = rng.uniform(-1.0, 1.0, a.T)
        t0 = time.perf_counter()
        ns["generate_trajectory_25D"](-10.0, -10.0, np.array([1.0, 0.0, 0.0]), v, w, 0.045, a.T, res, X, Y, Z)
        if k >= 2:
            times.append(time.perf_counter() - t0)
    per_step = float(np.median(times)) / a.T
    print(json.dumps({
        "what": "reference generate_trajectory_25D as written (displacement_on_surface.py:317-369), one core",
        "where": "build container CPU (not the GPU box: the reference tree does not travel)",
        "samples_timed": a.samples, "T": a.T, "us_per_sample_step": per_step * 1e6,
        "sample_steps_per_s_one_core": 1.0 / per_step,
        "C1_iteration_rollout_only_s": per_step * 1024 * 50, "cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
