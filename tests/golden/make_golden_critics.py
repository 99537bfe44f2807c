"""Generates tests/golden/reference_dormant_critics.npz: per-sample values of the reference's DORMANT critics --
`_path_orientation_critic` (critics_warp.py:44-83), `_avoid_slope` (:131-166) and `_goal_angle_critic` (:5-41), which
are defined in the reference but whose `costs[tid] +=` lines are commented out (:324, :326) or absent -- computed by
CALLING THE REFERENCE'S OWN FUNCTIONS (imported unmodified from /root/reference, executed under oracle/warp_shim.py)
on the K x T trajectories the reference's rollout kernel produced (tests/golden/reference_mppi_steps.npz).

Run in the build container only, after make_golden_warp.py:
    python tests/golden/make_golden_critics.py

It also runs the reference's `_evaluate_trajectories_kernel` with its two commented `costs[tid] +=` lines (:324
`_path_orientation_critic`, :326 `50.5*_avoid_slope`) RE-ENABLED -- the source text is read from /root/reference, the
two comment markers are removed in memory, and the result is executed under the shim -- which pins the position of the
two dormant terms in the cost sum (`reenabled_cost`).

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


def reenabled_evaluate_kernel(cw):
    """The reference's evaluation kernel with the comment markers of its lines 324 and 326 removed (in memory)."""
    import linecache
    import types
    path = cw.__file__
    lines = open(path).read().split("\n")
    for ln, what in ((324, "_path_orientation_critic"), (326, "50.5*_avoid_slope(")):
        text = lines[ln - 1]
        assert text.lstrip().startswith("# costs[tid] +=") and what in text, (ln, text)
        lines[ln - 1] = text.replace("# costs[tid]", "costs[tid]", 1)
    src = "\n".join(lines)
    name = path + " (:324, :326 re-enabled)"
    linecache.cache[name] = (len(src), None, src.splitlines(True), name)       # inspect.getsource (used by the shim)
    mod = types.ModuleType("critics_warp_reenabled")
    exec(compile(src, name, "exec"), mod.__dict__)
    return mod._evaluate_trajectories_kernel


def main():
    sys.path.insert(0, ROOT)
    from oracle import warp_shim as wp
    sys.modules["warp"] = wp
    sys.path.insert(0, REF)
    import thesis_master.warp_implementation.critics_warp as cw
    evaluate = reenabled_evaluate_kernel(cw)
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
            lw_wp = wp.array([wp.vec3f(r) for r in z[f"{pre}/out/lw"].reshape(K * T, 3)], dtype=wp.vec3f)
            rw_wp = wp.array([wp.vec3f(r) for r in z[f"{pre}/out/rw"].reshape(K * T, 3)], dtype=wp.vec3f)
            v_wp = wp.array(z[f"{pre}/out/v"].reshape(K * T), dtype=wp.float32)
            cm = z[f"{name}/costmap"]
            cm_wp = wp.array(cm.reshape(-1), dtype=wp.float32)
            hw, _, cres, _, _, _, horizon, _ = (float(v) for v in z[f"{name}/fmeta"])
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
                costs = wp.zeros(K, dtype=wp.float32)
                with np.errstate(divide="ignore", invalid="ignore"):
                    wp.launch(kernel=evaluate, dim=K,
                              inputs=[x, y, goal, theta, traj_wp, lw_wp, rw_wp, v_wp, 2.0, K, T, hw, cres,
                                      int(cm.shape[0]), cm_wp, horizon, costs])
                out[f"{pre}/{tag}/reenabled_cost"] = costs.numpy()
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
