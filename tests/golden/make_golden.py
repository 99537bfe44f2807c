"""Generates tests/golden/reference_functions.npz by EXECUTING the reference's own CPU projection functions
(/root/reference/thesis_master/python_mppi_projection/displacement_on_surface.py:48-466).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py

The script is a module-level program that imports matplotlib and plots, and calls the removed `np.int`
(:195-196); we exec only the function-definition block with `np.int(` -> `int(`.  No reference source is
copied into this repo: only the numeric outputs are stored.

What is pinned (the functions whose semantics coincide with the Warp kernels):
  normal_on_grid            <-> projection_warp.py:129-151
  get_heading_tangent_vector<-> projection_warp.py:168-190
  update_position           <-> projection_warp.py:207-248 (SciPy rotvec == Rodrigues)
  bilinear_interpolator     <-> projection_warp.py:70-100 for non-negative coordinates (floor == trunc there)
  generate_trajectory_2D    <-> projection_warp.py:353-382 (flat unicycle rollout)
"""
import os

import numpy as np

SRC = "/root/reference/thesis_master/python_mppi_projection/displacement_on_surface.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_functions.npz")


def load_reference_functions():
    lines = open(SRC).read().split("\n")
    block = "\n".join(lines[47:466]).replace("np.int(", "int(")
    ns = {}
    exec("import numpy as np\nfrom scipy.spatial.transform import Rotation as R\n" + block, ns)
    return ns


def main():
    ns = load_reference_functions()
    rng = np.random.default_rng(2024)
    N = 256
    res = 0.1
    q = rng.normal(0.0, 0.3, size=(N, 2, 2))
    normals = np.stack([ns["normal_on_grid"](q[i], res) for i in range(N)])
    heads = rng.normal(size=(N, 3))
    heads /= np.linalg.norm(heads, axis=1, keepdims=True)
    tangents = np.stack([ns["get_heading_tangent_vector"](normals[i], heads[i]) for i in range(N)])
    xy = rng.uniform(-15, 15, size=(N, 2))
    v = rng.uniform(0, 2, size=N)
    w = rng.uniform(-1, 1, size=N)
    dt = 0.045
    upd = np.zeros((N, 5))
    for i in range(N):
        nx, ny, nh = ns["update_position"](xy[i, 0], xy[i, 1], tangents[i].copy(), v[i], w[i], normals[i], dt)
        upd[i] = [nx, ny, *nh]
    xy_pos = rng.uniform(0.0, 15.0, size=(N, 2))
    bil = np.array([ns["bilinear_interpolator"](xy_pos[i, 0], xy_pos[i, 1], q[i], res) for i in range(N)])
    # flat 2-D rollout with time-varying commands
    T2 = 200
    v2 = rng.uniform(0.2, 2.0, size=T2)
    w2 = rng.uniform(-1.0, 1.0, size=T2)
    h2 = np.array([0.6, 0.8, 0.0])
    traj2 = ns["generate_trajectory_2D"](-3.0, 2.0, h2.copy(), v2, w2, 0.045, T2)
    # the reference's own saved 2-D trajectory (v = 1.5, w = 0, dt = 0.01 from (-14, -4), heading +x):
    # every 50th row of trajectory_2D.csv plus the last one -- the one golden file the reference ships.
    csv = np.loadtxt(os.path.join(os.path.dirname(SRC), "trajectory_2D.csv"), delimiter=",", skiprows=1)
    csv_idx = np.unique(np.concatenate([np.arange(0, csv.shape[0], 50), [csv.shape[0] - 1]]))
    np.savez_compressed(OUT, csv_idx=csv_idx, csv_rows=csv[csv_idx], csv_len=csv.shape[0], res=res, dt=dt, q=q, normals=normals, heads=heads, tangents=tangents, xy=xy, v=v, w=w,
                        upd=upd, xy_pos=xy_pos, bil=bil, v2=v2, w2=w2, h2=h2, traj2=traj2)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
