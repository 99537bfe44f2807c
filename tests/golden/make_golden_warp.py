"""Generates tests/golden/reference_mppi_steps.npz by RUNNING THE REFERENCE'S OWN CONTROLLER AND KERNEL SOURCES
(thesis_master/warp_implementation/{MPPI_isaac,sampling_warp,projection_warp,critics_warp}.py, imported unmodified
from /root/reference) under oracle/warp_shim.py, a pure-Python interpreter of the Warp scalar model.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_warp.py

What runs: the reference's `Surface` (crater DEM + cv2 distance-transform costmap), `Robot`, `MPPI_Controller` and
its `run()` closed loop -- warp_setup, reset, the nine `wp.launch` calls of MPPI_step, the host-side sigma / wheel
speed / pose feedback -- for a few iterations, on a small scene (K = 64 samples, T = 30 steps; the interpreter
executes one simulated thread at a time).  Only numeric inputs and outputs are stored; no reference source is copied.
The one substitution is wp.randn (third-party PCG + Box-Muller, absent): the shim returns a deterministic normal per
RNG state, and the same values are stored as eps1/eps2[k, t] using the reference's own state formula
(sampling_warp.py:71-92) so that the oracle and the CUDA path can be fed the identical noise.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "reference_mppi_steps.npz")

CONFIG = """
frame_work:
  robot_radius: 1.2
controller:
  number_of_iterations: {T}
  dt: 0.045
  number_of_trajectories: {K}
velocities:
  initial_linear_velocity: 0.0
  min_linear_velocity: 0.0
  max_linear_velocity: 2.0
  initial_angular_velocity: 0.00
  min_angular_velocity: -1.0
  max_angular_velocity: 1.0
inputs:
  std_dev_u1: 0.25
  std_dev_u2: 0.25
  min_u1: -1
  max_u1: 1
  min_u2: -1
  max_u2: 1
cost_evaluation:
  temperature: {lam}
"""


def install_shim():
    sys.path.insert(0, ROOT)
    from oracle import warp_shim
    sys.modules["warp"] = warp_shim
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    return warp_shim


def run_scenario(ref, wp, name, proj, K, T, lam, n_steps, start, heading, goal, wheel0=(0.0, 0.0)):
    gs, hw = 160, 8.0
    bumps = [((-1.5, 1.0), 1.4, 2.5), ((3.0, -1.0), 2.0, 3.0), ((0.5, 4.0), 0.9, 1.5), ((-4.0, -4.0), 1.2, 2.0)]
    obstacles = [(1.0, 2.0, 0.6), (-2.0, -1.0, 0.8), (3.5, 3.0, 0.5), (0.0, -3.0, 0.7), (-3.0, 3.5, 0.4)]
    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    cfg.write(CONFIG.format(K=K, T=T, lam=lam))
    cfg.close()
    surface = ref.Surface("manual", "", "manual", "", gs, hw, (0.0, 0.0), bumps, 0.3, obstacles)
    robot = ref.Robot(start[0], start[1], heading, cfg.name)
    robot.left_wheel_speed, robot.right_wheel_speed = wheel0
    ctrl = ref.MPPI_Controller(surface, robot, cfg.name, goal[0], goal[1], 2.2)

    steps = []
    seeds = []
    real_launch = wp.launch

    def spy_launch(kernel=None, dim=None, inputs=(), device=None, **kw):
        if kernel.__name__ == "_generate_inputs_kernel":
            seeds.append(int(inputs[1]))
        return real_launch(kernel=kernel, dim=dim, inputs=inputs, device=device, **kw)

    ref.wp.launch = spy_launch
    real_step = ctrl.MPPI_step

    def spy_step(proj):
        hv = np.asarray(robot.heading_vector, np.float64)
        pre = dict(x=robot.x[-1], y=robot.y[-1], heading=hv / np.linalg.norm(hv),
                   wheel_l=robot.left_wheel_speed, wheel_r=robot.right_wheel_speed,
                   sigma1=ctrl.std_dev_u1, sigma2=ctrl.std_dev_u2,
                   nominal1=ctrl.optimal_u1_wp.numpy(), nominal2=ctrl.optimal_u2_wp.numpy())
        real_step(proj=proj)
        seed = seeds[-1]
        tid = np.arange(K * T, dtype=np.int64)
        last = (tid % T) == (T - 1)
        s1 = np.where(last, seed + tid + 3 * T, seed + tid + T)          # sampling_warp.py:73,84
        s2 = np.where(last, seed + tid + 4 * T, seed + tid + 2 * T)      # sampling_warp.py:78,89
        eps1 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s1], np.float32).reshape(K, T)
        eps2 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s2], np.float32).reshape(K, T)
        post = dict(seed=seed, eps1=eps1, eps2=eps2,
                    u1=ctrl.u1.numpy().reshape(K, T), u2=ctrl.u2.numpy().reshape(K, T),
                    v=ctrl.linear_velocities.numpy().reshape(K, T), w=ctrl.angular_velocities.numpy().reshape(K, T),
                    traj=ctrl.trajectories.numpy().reshape(K, T, 3),
                    heading_vectors=ctrl.heading_vectors.numpy().reshape(K, T, 3),
                    lw=ctrl.left_wheel_pos.numpy().reshape(K, T, 3), rw=ctrl.right_wheel_pos.numpy().reshape(K, T, 3),
                    costs=ctrl.costs_wp.numpy(), weights=ctrl.weights_wp.numpy(),
                    min_cost=ctrl.min_cost.numpy()[0], weights_sum=ctrl.weights_sum.numpy()[0],
                    out_nominal1=ctrl.optimal_u1_wp.numpy(), out_nominal2=ctrl.optimal_u2_wp.numpy(),
                    opt_v=ctrl.optimal_lin_vel_wp.numpy(), opt_w=ctrl.optimal_ang_vel_wp.numpy(),
                    sim_traj=ctrl.trajectories_sim.numpy(), sim_heading=ctrl.heading_vectors_sim.numpy())
        steps.append((pre, post))

    ctrl.MPPI_step = spy_step
    ctrl.loop = 3500 - n_steps                  # the reference's own run() loop executes exactly n_steps iterations
    ctrl.run(proj)
    ref.wp.launch = real_launch
    os.unlink(cfg.name)

    out = {f"{name}/Z": np.asarray(surface.Z, np.float32), f"{name}/costmap": np.asarray(surface.costmap, np.float32),
           f"{name}/meta": np.array([K, T, n_steps, gs, surface.costmap_size, 3 if proj == "3d" else 2], np.int64),
           f"{name}/fmeta": np.array([hw, surface.resolution, surface.costmap_resolution, lam, goal[0], goal[1],
                                      ctrl.horizon, robot.radius], np.float64),
           f"{name}/final_pose": np.array([robot.x[-1], robot.y[-1], robot.z[-1], *np.asarray(robot.heading_vector)],
                                          np.float64)}
    for i, (pre, post) in enumerate(steps):
        for k, v in pre.items():
            out[f"{name}/step{i}/in/{k}"] = np.asarray(v)
        for k, v in post.items():
            out[f"{name}/step{i}/out/{k}"] = np.asarray(v)
    return out


def run_velocity_space(ref, wp, name, K, T, lam, n_steps, start, heading, goal):
    """Velocity-space MPPI (the unicycle input model): the reference's `_generate_velocities_kernel`
    (sampling_warp.py:10-48) feeding today's rollout / critic / update kernels.  MPPI_isaac.py never launches that
    kernel (its driver was old_files/run_mppi.py, whose other launches are stale), so THIS script issues the launches,
    in the order of MPPI_step (MPPI_isaac.py:578-670, 696-720) with (v, w) in place of the filtered wheel inputs; all
    arithmetic still runs inside the reference's own kernel sources."""
    gs, hw = 160, 8.0
    bumps = [((-1.5, 1.0), 1.4, 2.5), ((3.0, -1.0), 2.0, 3.0), ((0.5, 4.0), 0.9, 1.5), ((-4.0, -4.0), 1.2, 2.0)]
    obstacles = [(1.0, 2.0, 0.6), (-2.0, -1.0, 0.8), (3.5, 3.0, 0.5), (0.0, -3.0, 0.7), (-3.0, 3.5, 0.4)]
    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    cfg.write(CONFIG.format(K=K, T=T, lam=lam))
    cfg.close()
    surface = ref.Surface("manual", "", "manual", "", gs, hw, (0.0, 0.0), bumps, 0.3, obstacles)
    robot = ref.Robot(start[0], start[1], heading, cfg.name)
    c = ref.MPPI_Controller(surface, robot, cfg.name, goal[0], goal[1], 2.2)
    c.warp_setup()
    os.unlink(cfg.name)
    std_lin, std_ang = 0.3, 0.2
    rng = np.random.default_rng(11)
    out = {f"{name}/Z": np.asarray(surface.Z, np.float32), f"{name}/costmap": np.asarray(surface.costmap, np.float32),
           f"{name}/meta": np.array([K, T, n_steps, gs, surface.costmap_size, 3], np.int64),
           f"{name}/fmeta": np.array([hw, surface.resolution, surface.costmap_resolution, lam, goal[0], goal[1],
                                      c.horizon, robot.radius], np.float64)}
    for i in range(n_steps):
        c.reset("controller")
        hv = np.asarray(robot.heading_vector, np.float64)
        pre = dict(x=robot.x[-1], y=robot.y[-1], heading=hv / np.linalg.norm(hv), wheel_l=0.0, wheel_r=0.0,
                   sigma1=std_lin, sigma2=std_ang,
                   nominal1=c.optimal_lin_vel_wp.numpy(), nominal2=c.optimal_ang_vel_wp.numpy())
        seed = int(rng.integers(T + 1, 1000))
        wp.launch(ref._generate_velocities_kernel, dim=K * T,
                  inputs=[T, seed, c.optimal_lin_vel_wp, c.optimal_ang_vel_wp, std_lin, c.v_min_linear, c.v_max_linear,
                          std_ang, c.v_min_angular, c.v_max_angular, c.linear_velocities, c.angular_velocities])
        wp.launch(ref._generate_trajectories_kernel, dim=K,
                  inputs=[c.position, -surface.half_width, -surface.half_width, surface.grid_size, c.q,
                          surface.resolution, c.Z_wp, c.height, c.normal, c.heading_vectors, c.previous_heading_vector,
                          T, c.linear_velocities, c.angular_velocities, c.dt, c.trajectories, c.left_wheel_pos,
                          c.right_wheel_pos])
        wp.launch(kernel=ref._evaluate_trajectories_kernel, dim=K,
                  inputs=[robot.x[-1], robot.y[-1], c.goal, c.goal_orientation, c.trajectories, c.left_wheel_pos,
                          c.right_wheel_pos, c.linear_velocities, c.v_max_linear, K, T, surface.half_width,
                          surface.costmap_resolution, surface.costmap_size, c.costmap_wp, c.horizon, c.costs_wp])
        wp.launch(kernel=ref._compute_weights, dim=K, inputs=[c.costs_wp, c.min_cost, c.weights_wp, c.temperature])
        wp.launch(kernel=ref._compute_sum, dim=K, inputs=[c.weights_wp, c.weights_sum])
        c.optimal_lin_vel_wp.zero_()
        c.optimal_ang_vel_wp.zero_()
        wp.launch(kernel=ref._compute_weighted_sum, dim=K,
                  inputs=[c.weights_wp, T, c.linear_velocities, c.angular_velocities, c.weights_sum,
                          c.optimal_lin_vel_wp, c.optimal_ang_vel_wp])
        c.reset("sim")                                   # also zeroes optimal_u1/u2 (unused in this mode)
        wp.launch(kernel=ref._generate_trajectories_kernel, dim=1,
                  inputs=[c.position_sim, -surface.half_width, -surface.half_width, surface.grid_size, c.q_sim,
                          surface.resolution, c.Z_wp, c.height_sim, c.normal_sim, c.heading_vectors_sim,
                          c.previous_heading_vector, T, c.optimal_lin_vel_wp, c.optimal_ang_vel_wp, c.dt,
                          c.trajectories_sim, c.left_wheel_pos_sim, c.right_wheel_pos_sim])
        tid = np.arange(K * T, dtype=np.int64)
        eps1 = np.array([wp.randn_from_state(wp.uint32(s_)) for s_ in seed + tid], np.float32).reshape(K, T)
        eps2 = np.array([wp.randn_from_state(wp.uint32(s_)) for s_ in seed + tid + T], np.float32).reshape(K, T)
        post = dict(seed=seed, eps1=eps1, eps2=eps2,
                    u1=c.linear_velocities.numpy().reshape(K, T), u2=c.angular_velocities.numpy().reshape(K, T),
                    v=c.linear_velocities.numpy().reshape(K, T), w=c.angular_velocities.numpy().reshape(K, T),
                    traj=c.trajectories.numpy().reshape(K, T, 3),
                    heading_vectors=c.heading_vectors.numpy().reshape(K, T, 3),
                    lw=c.left_wheel_pos.numpy().reshape(K, T, 3), rw=c.right_wheel_pos.numpy().reshape(K, T, 3),
                    costs=c.costs_wp.numpy(), weights=c.weights_wp.numpy(),
                    min_cost=c.min_cost.numpy()[0], weights_sum=c.weights_sum.numpy()[0],
                    out_nominal1=c.optimal_lin_vel_wp.numpy(), out_nominal2=c.optimal_ang_vel_wp.numpy(),
                    opt_v=c.optimal_lin_vel_wp.numpy(), opt_w=c.optimal_ang_vel_wp.numpy(),
                    sim_traj=c.trajectories_sim.numpy(), sim_heading=c.heading_vectors_sim.numpy())
        for k_, v_ in pre.items():
            out[f"{name}/step{i}/in/{k_}"] = np.asarray(v_)
        for k_, v_ in post.items():
            out[f"{name}/step{i}/out/{k_}"] = np.asarray(v_)
        t0, h0 = c.trajectories_sim.numpy()[0], c.heading_vectors_sim.numpy()[0]
        robot.update_position(t0[0], t0[1], t0[2], h0)
    return out


def main():
    wp = install_shim()
    import thesis_master.warp_implementation.MPPI_isaac as ref
    out = {}
    # A: 3-D projection, reference temperature (softmax ~ argmin), three closed-loop iterations of run()
    out.update(run_scenario(ref, wp, "A3d", "3d", K=64, T=30, lam=0.3, n_steps=3,
                            start=(-2.31, -2.87), heading=(1.0, 0.35, 0.0), goal=(5.1, 4.3), wheel0=(0.3, 0.5)))
    # B: 3-D, large temperature so that many samples carry weight (the weighted update is really exercised);
    #    goal inside the horizon -> near-goal branch of the path critic
    out.update(run_scenario(ref, wp, "B3d_hot", "3d", K=48, T=24, lam=50000.0, n_steps=2,
                            start=(1.26, -0.63), heading=(-0.4, 1.0, 0.0), goal=(1.9, 0.9), wheel0=(0.8, 0.6)))
    # C: flat 2-D projection mode
    out.update(run_scenario(ref, wp, "C2d", "2d", K=48, T=24, lam=0.3, n_steps=2,
                            start=(-3.07, 2.18), heading=(0.2, -1.0, 0.0), goal=(4.0, -5.0)))
    # D: velocity-space (unicycle) input model, sampling_warp.py:10-48
    out.update(run_velocity_space(ref, wp, "D3d_unicycle", K=48, T=24, lam=2000.0, n_steps=2,
                                  start=(2.2, 3.1), heading=(-1.0, -0.2, 0.0), goal=(-5.0, -4.0)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
