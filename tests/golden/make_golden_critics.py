"""Generates tests/golden/reference_dormant_critics.npz: per-sample values of the reference's DORMANT critics --
`_path_orientation_critic` (critics_warp.py:44-83), `_avoid_slope` (:131-166) and `_goal_angle_critic` (:5-41), which
are defined in the reference but whose `costs[tid] +=` lines are commented out (:324, :326) or absent -- computed by
CALLING THE REFERENCE'S OWN FUNCTIONS (imported unmodified from /root/reference, executed under oracle/warp_shim.py)
on the K x T trajectories the reference's rollout kernel produced (tests/golden/reference_mppi_steps.npz).

Run in the build container only, after make_golden_warp.py:
    python tests/golden/make_golden_critics.py

Each scenario step is evaluated for three goals: the scenario's own, one 0.3 m from the robot (inside the 0.5 m radius
that arms the goal-angle critic) and one behind the robot (so that the orientation critic is non-zero for most
samples).  Only numbers are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SRC = os.path.join(HERE, "reference_mppi_steps.npz")
OUT = os.path.join(HERE, "reference_dormant_critics.npz")


def main():
    sys.path.insert(0, ROOT)
    from oracle import warp_shim as wp
    sys.modules["warp"] = wp
    sys.path.insert(0, REF)
    import thesis_master.warp_implementation.critics_warp as cw
    z = np.load(SRC)
    names = sorted({k.split("/")[0] for k in z.files})
    out = {}
    for name in names:
        K, T, n_steps = (int(v) for v in z[f"{name}/meta"][:3])
        gx, gy = (float(v) for v in z[f"{name}/fmeta"][4:6])
        for i in range(n_steps):
            pre = f"{name}/step{i}"
            x, y = float(z[f"{pre}/in/x"]), float(z[f"{pre}/in/y"])
            hd = np.asarray(z[f"{pre}/in/heading"], np.float64)
            traj = z[f"{pre}/out/traj"].reshape(K * T, 3)
            traj_wp = wp.array([wp.vec3f(r) for r in traj], dtype=wp.vec3f)
            goals = {"own": (gx, gy),
                     "near": (x + 0.3 * hd[0] / np.hypot(hd[0], hd[1]), y + 0.3 * hd[1] / np.hypot(hd[0], hd[1])),
                     "behind": (x - 3.0 * hd[0] + 0.4, y - 3.0 * hd[1] - 0.2)}
            theta = 2.2 if i % 2 == 0 else -0.7
            out[f"{pre}/goal_theta"] = np.float32(theta)
            slope = np.zeros(K, np.float32)
            for k in range(K):
                slope[k] = cw._avoid_slope(traj_wp, wp.float(k) * wp.float(T), wp.float(T))
            out[f"{pre}/slope_path"] = slope
            for tag, g in goals.items():
                goal = wp.vec2f(g[0], g[1])
                orient = np.zeros(K, np.float32)
                angle = np.zeros(K, np.float32)
                with np.errstate(divide="ignore", invalid="ignore"):
                    for k in range(K):
                        s0 = wp.float(k) * wp.float(T)
                        orient[k] = cw._path_orientation_critic(wp.float(x), wp.float(y), goal, traj_wp, s0, wp.float(T))
                        angle[k] = cw._goal_angle_critic(wp.float(x), wp.float(y), goal, wp.float(theta), traj_wp, s0,
                                                         wp.float(T))
                out[f"{pre}/{tag}/goal"] = np.array(g, np.float64)
                out[f"{pre}/{tag}/orient"] = orient
                out[f"{pre}/{tag}/goal_angle"] = angle
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")
    for k in sorted(out):
        if k.endswith(("orient", "goal_angle", "slope_path")):
            v = out[k]
            print(f"  {k:40s} nonzero {int(np.count_nonzero(v)):3d}/{v.size}  max {float(np.nanmax(v)):.5g}")


if __name__ == "__main__":
    main()
