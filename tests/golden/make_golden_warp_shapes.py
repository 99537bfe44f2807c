"""Generates tests/golden/reference_mppi_baseline_shapes.npz: ONE `MPPI_Controller.MPPI_step` of the reference
(thesis_master/warp_implementation/MPPI_isaac.py:505-720 and the three kernel files, imported UNMODIFIED from
/root/reference, interpreted by oracle/warp_shim.py) at the shapes BASELINE.json names:

  C1      K = 1024, T = 50,  1500^2 DEM (0.1 m), 750^2 costmap (0.2 m), the bench start (-60.57, -60.23)
  C1rock  same shape, started inside the rock field on a crater wall with a warm nominal (lethal cells, slopes)
  C2      K = 4096, T = 100, the bench start
  C2rock  K = 4096, T = 100, the rock-field start

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_warp_shapes.py            # ~6 minutes, one process

The scene is the reference's own `Surface("manual", ...)` with the nine craters of MPPI_OO_current.py:730-740; the
750^2 costmap of the bench recipe is assigned to `surface.costmap / costmap_size / costmap_resolution`, the attributes
the Isaac driver itself rewrites between steps (visual_terrain_stack_full_terrain.py:561-563) -- the reference's own
constructor would give int(1500 / 8) = 187 cells (MPPI_isaac.py:271), BASELINE.json names 750.

To keep the fixture small only what a test cannot regenerate is stored:
  * the DEM / costmap WINDOW the rollouts can reach (the tests paste it into NaN-filled full-size maps, so a read
    outside the window poisons the result instead of passing silently);
  * the RNG seed -- eps1 / eps2 are `warp_shim.randn_from_state` of the reference's own state formula
    (sampling_warp.py:71-92), recomputed by the tests; a float64 checksum of both arrays is stored;
  * per-sample outputs in full (costs, weights), K x T arrays for `KEEP` evenly strided samples.
No reference source is copied; only numeric inputs and outputs are stored.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "reference_mppi_baseline_shapes.npz")
sys.path.insert(0, HERE)
from make_golden_warp import CONFIG, install_shim          # noqa: E402

KEEP = 128              # samples (evenly strided) that keep their K x T arrays
WINDOW_M = 12.0         # half side of the stored terrain window (reach at T = 100 is 9.2 m)


def eps_from_seed(wp, seed, K, T):
    """The noise the shim's wp.randn hands the reference's _generate_inputs_kernel (sampling_warp.py:71-92)."""
    tid = np.arange(K * T, dtype=np.int64)
    last = (tid % T) == (T - 1)
    s1 = np.where(last, seed + tid + 3 * T, seed + tid + T)
    s2 = np.where(last, seed + tid + 4 * T, seed + tid + 2 * T)
    e1 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s1], np.float32).reshape(K, T)
    e2 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s2], np.float32).reshape(K, T)
    return e1, e2


def window(gs, hw, res, x, y, half):
    """[j0:j1, i0:i1] of a [gs, gs] map (row 0 at y = +hw) covering the square of half side `half` around (x, y)."""
    i0, i1 = int((x - half + hw) / res), int((x + half + hw) / res) + 2
    j0, j1 = int((hw - y - half) / res), int((hw - y + half) / res) + 2
    return max(i0, 0), min(i1, gs), max(j0, 0), min(j1, gs)


def run_shape(ref, wp, name, K, T, start, heading, goal, wheel0=(0.0, 0.0), nominal=None, sigma=None, keep=KEEP):
    from mppi_b200 import synthetic as syn
    gs, hw, cms = 1500, 75.0, 750
    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    cfg.write(CONFIG.format(K=K, T=T, lam=0.3))
    cfg.close()
    surface = ref.Surface("manual", "", "none", "", gs, hw, (0.0, 0.0), syn.NINE_CRATERS, 0.3, [])
    surface.costmap = syn.rock_costmap(cms, hw)
    surface.costmap_size = cms
    surface.costmap_resolution = 2 * hw / cms
    robot = ref.Robot(start[0], start[1], heading, cfg.name)
    robot.left_wheel_speed, robot.right_wheel_speed = wheel0
    ctrl = ref.MPPI_Controller(surface, robot, cfg.name, goal[0], goal[1], 2.2)
    os.unlink(cfg.name)
    ctrl.warp_setup()
    ctrl.reset("controller")
    if nominal is not None:
        ctrl.optimal_u1_wp.assign(np.asarray(nominal[0], np.float32))
        ctrl.optimal_u2_wp.assign(np.asarray(nominal[1], np.float32))
    if sigma is not None:
        ctrl.std_dev_u1, ctrl.std_dev_u2 = sigma
    seeds = []
    real_launch = wp.launch

    def spy_launch(kernel=None, dim=None, inputs=(), device=None, **kw):
        if kernel.__name__ == "_generate_inputs_kernel":
            seeds.append(int(inputs[1]))
        return real_launch(kernel=kernel, dim=dim, inputs=inputs, device=device, **kw)

    ref.wp.launch = spy_launch
    hv = np.asarray(robot.heading_vector, np.float64)
    pre = dict(x=robot.x[-1], y=robot.y[-1], heading=hv / np.linalg.norm(hv), wheel_l=robot.left_wheel_speed,
               wheel_r=robot.right_wheel_speed, sigma1=ctrl.std_dev_u1, sigma2=ctrl.std_dev_u2,
               nominal1=ctrl.optimal_u1_wp.numpy().copy(), nominal2=ctrl.optimal_u2_wp.numpy().copy())
    t0 = time.time()
    ctrl.MPPI_step(proj="3d")
    ref.wp.launch = real_launch
    seed = seeds[-1]
    e1, e2 = eps_from_seed(wp, seed, K, T)
    u1 = ctrl.u1.numpy().reshape(K, T)
    # the stored checksum must describe the noise the kernels really consumed: u = clamp(nominal_shifted + sigma eps)
    shifted = np.concatenate([pre["nominal1"][1:], pre["nominal1"][-1:]])
    assert np.array_equal(u1, np.clip(shifted[None, :] + np.float32(pre["sigma1"]) * e1, -1, 1).astype(np.float32))
    Z = np.asarray(surface.Z, np.float32)
    cm = np.asarray(surface.costmap, np.float32)
    di0, di1, dj0, dj1 = window(gs, hw, surface.resolution, start[0], start[1], WINDOW_M)
    ci0, ci1, cj0, cj1 = window(cms, hw, surface.costmap_resolution, start[0], start[1], WINDOW_M)
    SUB = K // keep
    sub = slice(0, K, SUB)
    r3 = lambda a: a.numpy().reshape(K, T, 3)[sub]                     # noqa: E731
    r1 = lambda a: a.numpy().reshape(K, T)[sub]                        # noqa: E731
    out = {
        "meta": np.array([K, T, gs, cms, SUB, seed, di0, di1, dj0, dj1, ci0, ci1, cj0, cj1], np.int64),
        "fmeta": np.array([hw, surface.resolution, surface.costmap_resolution, 0.3, goal[0], goal[1], ctrl.horizon,
                           robot.radius, float(e1.astype(np.float64).sum()), float(e2.astype(np.float64).sum())],
                          np.float64),
        "Z_window": Z[dj0:dj1, di0:di1], "costmap_window": cm[cj0:cj1, ci0:ci1],
        "out/u1": r1(ctrl.u1), "out/u2": r1(ctrl.u2), "out/v": r1(ctrl.linear_velocities),
        "out/w": r1(ctrl.angular_velocities), "out/traj": r3(ctrl.trajectories),
        "out/heading_vectors": r3(ctrl.heading_vectors), "out/lw": r3(ctrl.left_wheel_pos),
        "out/rw": r3(ctrl.right_wheel_pos), "out/costs": ctrl.costs_wp.numpy(), "out/weights": ctrl.weights_wp.numpy(),
        "out/min_cost": ctrl.min_cost.numpy()[0], "out/weights_sum": ctrl.weights_sum.numpy()[0],
        "out/out_nominal1": ctrl.optimal_u1_wp.numpy(), "out/out_nominal2": ctrl.optimal_u2_wp.numpy(),
        "out/opt_v": ctrl.optimal_lin_vel_wp.numpy(), "out/opt_w": ctrl.optimal_ang_vel_wp.numpy(),
        "out/sim_traj": ctrl.trajectories_sim.numpy(), "out/sim_heading": ctrl.heading_vectors_sim.numpy(),
    }
    for k, v in pre.items():
        out["in/" + k] = np.asarray(v)
    costs = out["out/costs"]
    tr = ctrl.trajectories.numpy().reshape(K, T, 3)
    ix = ((tr[..., 0] + hw) / surface.costmap_resolution).astype(np.int64)        # critics_warp.py:245-248
    iy = ((-tr[..., 1] + hw) / surface.costmap_resolution).astype(np.int64)
    lethal = (cm[iy, ix] > 0.99).any(axis=1).sum()
    print(f"{name}: K={K} T={T} seed={seed} step {time.time() - t0:.0f} s, min cost {costs.min():.3f} at "
          f"{int(costs.argmin())}, samples crossing a lethal cell {lethal}, weights > 0: "
          f"{(out['out/weights'] > 0).sum()}, height range {tr[..., 2].min():.2f}..{tr[..., 2].max():.2f} m")
    return {f"{name}/{k}": np.asarray(v) for k, v in out.items()}


def main():
    wp = install_shim()
    import thesis_master.warp_implementation.MPPI_isaac as ref
    out = {}
    out.update(run_shape(ref, wp, "C1", 1024, 50, (-60.57, -60.23), (1.0, 0.0, 0.0), (65.80, 65.40)))
    # inside the rock field (rocks live within +-50 m), on the wall of the crater at (-20.67, -40.12), moving: the
    # rollouts cross lethal cells and real slopes, the nominal is warm and the two sigmas differ (run() adapts them)
    T = 50
    warm = (np.linspace(0.9, 0.5, T), np.linspace(0.6, 0.9, T))
    out.update(run_shape(ref, wp, "C1rock", 1024, T, (-16.37, -31.73), (0.6, -0.8, 0.0), (20.0, -48.0),
                         wheel0=(1.1, 1.4), nominal=warm, sigma=(0.4, 0.46)))
    out.update(run_shape(ref, wp, "C2", 4096, 100, (-60.57, -60.23), (1.0, 0.0, 0.0), (65.80, 65.40), keep=64))
    T = 100
    warm = (np.linspace(0.9, 0.5, T), np.linspace(0.6, 0.9, T))
    out.update(run_shape(ref, wp, "C2rock", 4096, T, (-16.37, -31.73), (0.6, -0.8, 0.0), (20.0, -48.0),
                         wheel0=(1.1, 1.4), nominal=warm, sigma=(0.4, 0.46), keep=64))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
