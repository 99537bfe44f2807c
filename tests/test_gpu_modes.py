"""GPU: the other operating modes of the core, each against the oracle --
sample-sharded partial/combine (ranks emulated on one GPU), multi-rover batches, the reference-shaped
`MPPI_Controller` facade (drop-in surface), validation/visualiser dumps."""
import ctypes as C

import numpy as np
import pytest

from util import default_state, terrain

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def close(a, b, floor=1e-3):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def make_core(K, T, name="C1", **kw):
    import torch
    from mppi_b200.core import Core
    dem, cm, hw = terrain(name)
    core = Core(K, T, **kw)
    core.set_terrain(torch.from_numpy(dem).cuda(), hw, torch.from_numpy(cm).cuda())
    return core, dem, cm, hw


def state_struct(st):
    from mppi_b200 import capi
    s = capi.MppiState()
    for k, v in st.items():
        setattr(s, k, float(v))
    return s


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("lam", [0.3, 2000.0])
def test_sample_sharded_partials_fold_to_the_unsharded_update(oracle, G, lam):
    """G ranks each roll out K/G samples keyed by GLOBAL sample id, publish their softmax
    partial, and the rank-ordered combine reproduces the single-GPU update (argmin exactly, nominal <= 1e-6 rel).
    Ranks are emulated as G sequential launches on one GPU -- the all-gather itself is covered by the gloo test."""
    import torch
    K, T = 4096, 60
    st = default_state()
    ref_eps = oracle.philox_normals(5, 11, K, T)
    dem, cm, hw = terrain("C1")
    nom = np.full(T, 0.35, np.float32)
    ref = oracle.mppi_step(oracle.make_params(K=K, T=T, lam=lam), dem, hw, cm, st, nom, nom, ref_eps[0], ref_eps[1],
                           dump=["cost"])
    core, *_ = make_core(K // G, T, lambda_=lam)
    s = state_struct(st)
    parts = torch.zeros((G, core.partial_floats()), device="cuda")
    costs = []
    for g in range(G):
        core.set_nominal(nom, nom)
        core.step_partial(s, parts[g], k_begin=g * (K // G), seed=5, offset=11)
        torch.cuda.synchronize()
        costs.append(core.costs[0].cpu().numpy().copy())
    assert np.array_equal(np.concatenate(costs), ref.dump["cost"])          # costs independent of the sharding
    core.set_nominal(nom, nom)
    core.combine_partials(s, parts, G)
    torch.cuda.synchronize()
    got = core.read_stats()
    assert got["argmin"] == ref.argmin and got["min_cost"] == ref.min_cost
    assert close(core.optimal_u1[0].cpu().numpy(), ref.nominal1_f64) < 1e-5
    assert close(core.optimal_u2[0].cpu().numpy(), ref.nominal2_f64) < 1e-5
    assert close(got["weights_sum"], ref.weights_sum) < 1e-5
    # the GPU partial layout is what the oracle's combine expects: fold it on the CPU too
    p = parts.cpu().numpy()
    n1, n2, m, arg, S = oracle.combine_partials(np.concatenate([p[:, :3], p[:, 4:]], axis=1), T, lam)
    assert arg == ref.argmin and close(n1, ref.nominal1_f64) < 1e-5
    core.close()


def test_unsharded_and_sharded_agree_on_the_device(oracle):
    """Same controller, once with mppi_step and once with step_partial + combine (G = 1): identical outputs."""
    import torch
    K, T = 2048, 50
    st = state_struct(default_state())
    core, *_ = make_core(K, T, lambda_=50.0)
    core.step(st, seed=3, offset=9)
    torch.cuda.synchronize()
    a = (core.optimal_u1.cpu().numpy().copy(), core.optimal_v.cpu().numpy().copy(), core.read_stats())
    core.set_nominal(np.zeros(T, np.float32), np.zeros(T, np.float32))
    part = torch.zeros(core.partial_floats(), device="cuda")
    core.step_partial(st, part, k_begin=0, seed=3, offset=9)
    core.combine_partials(st, part, 1)
    torch.cuda.synchronize()
    b = (core.optimal_u1.cpu().numpy(), core.optimal_v.cpu().numpy(), core.read_stats())
    assert a[2]["argmin"] == b[2]["argmin"]
    assert close(b[0], a[0]) < 1e-6 and close(b[1], a[1]) < 1e-6
    core.close()


@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_multi_rover_batch_matches_per_rover_oracle(oracle, variant):
    """BASELINE config 4 in small: R independent rovers, own map / pose / goal / nominal each, one launch."""
    import torch
    from mppi_b200.core import Core
    from mppi_b200 import synthetic as syn
    R, K, T = 6, 256, 40
    hw, gs, cms = 12.8, 256, 128
    rng = np.random.default_rng(7)
    dems, cms_, states, noms = [], [], [], []
    for r in range(R):
        bumps = [((float(rng.uniform(-8, 8)), float(rng.uniform(-8, 8))), float(rng.uniform(0.5, 2.0)),
                  float(rng.uniform(1.0, 3.0))) for _ in range(5)]
        dems.append(syn.crater_dem(gs, hw, bumps=bumps).numpy())
        cms_.append(syn.rock_costmap(cms, hw, n_rocks=40, seed=100 + r))
        th = rng.uniform(0, 2 * np.pi)
        states.append(default_state(x=float(rng.uniform(-5, 5)), y=float(rng.uniform(-5, 5)), hx=float(np.cos(th)),
                                    hy=float(np.sin(th)), goal_x=float(rng.uniform(-9, 9)),
                                    goal_y=float(rng.uniform(-9, 9)), wheel_l=float(rng.uniform(0, 1)),
                                    wheel_r=float(rng.uniform(0, 1))))
        noms.append((rng.uniform(-0.5, 1, T).astype(np.float32), rng.uniform(-0.5, 1, T).astype(np.float32)))
    core = Core(K, T, max_rovers=R, variant=variant)
    core.set_terrain_batched(torch.from_numpy(np.stack(dems)).cuda(), hw, torch.from_numpy(np.stack(cms_)).cuda())
    core.set_nominal(np.stack([n[0] for n in noms]), np.stack([n[1] for n in noms]), n_rovers=R)
    sdev = Core.pack_states([state_struct(s) for s in states], core.device)
    core.step_batched(sdev, R, seed=21, offset=4)
    torch.cuda.synchronize()
    for r in range(R):
        e1, e2 = oracle.philox_normals(21, 4, K, T, rover=r)
        ref = oracle.mppi_step(oracle.make_params(K=K, T=T), dems[r], hw, cms_[r], states[r], noms[r][0], noms[r][1],
                               e1, e2, dump=["cost"])
        st = core.read_stats(r)
        assert np.array_equal(core.costs[r].cpu().numpy(), ref.dump["cost"]), r
        assert st["argmin"] == ref.argmin and st["oob"] == ref.oob_clamps == 0
        assert close(core.optimal_u1[r].cpu().numpy(), ref.nominal1) < RTOL
        assert close(core.optimal_v[r].cpu().numpy(), ref.opt_v) < RTOL
    core.close()


def test_controller_facade_is_a_drop_in(oracle):
    """The reference driver's call sequence (visual_terrain_stack_full_terrain.py:449-515) against the facade:
    Surface / Robot / MPPI_Controller construction, warp_setup, reset, MPPI_step, `.numpy()[0]` reads, attribute
    writes between steps, costmap assign, Z_wp swap, sim rollout reads -- and the numbers equal the oracle's."""
    import torch
    import mppi_b200
    from mppi_b200 import MPPI_Controller, Robot, Surface
    dem, cm, hw = terrain("C1")
    surface = Surface("", "", "", "", grid_size=1500, half_width=75.0, origin=(0, 0), bumps=[], radius_robot=0.3)
    surface.Z = dem
    surface.costmap_size, surface.costmap_resolution, surface.costmap = 750, 150.0 / 750, cm
    robot = Robot(x=-60.57, y=-60.23, heading_vector=[1.0, 0.0, 0.0], config_file=mppi_b200.DEFAULT_CONFIG)
    ctl = MPPI_Controller(surface, robot, mppi_b200.DEFAULT_CONFIG, goal_x=65.8, goal_y=65.4, goal_orientation=2.2,
                          seed=77)
    ctl.warp_setup()
    K, T = ctl.number_of_trajectories, ctl.number_of_iterations
    assert (K, T) == (1000, 100)
    n1 = np.zeros(T, np.float32)
    n2 = np.zeros(T, np.float32)
    st = default_state()
    p = oracle.make_params(K=K, T=T)
    for it in range(3):
        ctl.reset("controller")
        ctl.MPPI_step(proj="3d")
        lin = ctl.optimal_lin_vel_wp.numpy()[0]
        ang = ctl.optimal_ang_vel_wp.numpy()[0]
        e1, e2 = oracle.philox_normals(77, it, K, T)
        ref = oracle.mppi_step(p, dem, hw, cm, st, n1, n2, e1, e2, dump=["cost"])
        assert np.array_equal(ctl.costs_wp.numpy(), ref.dump["cost"])
        assert close([lin, ang], [ref.opt_v[0], ref.opt_w[0]]) < RTOL
        assert close(ctl.optimal_u1_wp.numpy(), ref.nominal1) < RTOL
        assert close(ctl.trajectories_sim.numpy(), ref.sim_traj, floor=1.0) < RTOL     # lazy launch 9
        assert close(ctl.heading_vectors_sim.numpy(), ref.sim_heading, floor=1.0) < RTOL
        assert ctl.stats()["argmin"] == ref.argmin
        # what the driver mutates between steps
        n1, n2 = ctl.optimal_u1_wp.numpy().copy(), ctl.optimal_u2_wp.numpy().copy()
        t0, h0 = ref.sim_traj[0], ref.sim_heading[0]
        robot.update_position(float(t0[0]), float(t0[1]), float(t0[2]), h0.astype(np.float64))
        ctl.std_dev_u1 = max(0.25, 0.25 - float(ang) ** 2 / 3)
        ctl.std_dev_u2 = max(0.25, 0.25 + float(ang) ** 2 / 3)
        robot.left_wheel_speed = float(lin) - float(ang) * robot.radius / 2
        robot.right_wheel_speed = float(lin) + float(ang) * robot.radius / 2
        hv = h0.astype(np.float64) / np.linalg.norm(h0.astype(np.float64))
        st = default_state(x=float(t0[0]), y=float(t0[1]), hx=hv[0], hy=hv[1], hz=hv[2],
                           wheel_l=robot.left_wheel_speed, wheel_r=robot.right_wheel_speed,
                           sigma1=ctl.std_dev_u1, sigma2=ctl.std_dev_u2)
    # K x T trajectories only on request (visualiser path)
    tr = ctl.trajectories.numpy()
    assert tr.shape == (K * T, 3)
    # block change: new costmap through assign(), DEM re-pointed zero-copy at another device array, goal moved
    cm2 = np.ascontiguousarray(cm[::-1]).copy()
    ctl.surface.costmap = cm2
    ctl.costmap_wp.assign(cm2.flatten())
    dem2 = torch.from_numpy(np.ascontiguousarray(dem.T)).cuda()
    ctl.Z_wp = dem2
    ctl.goal_x -= 3.0
    ctl.goal_y += 2.0
    ctl.MPPI_step(proj="3d")
    e1, e2 = oracle.philox_normals(77, 3, K, T)
    st["goal_x"], st["goal_y"] = ctl.goal_x, ctl.goal_y
    ref = oracle.mppi_step(p, np.ascontiguousarray(dem.T), hw, cm2, st, n1, n2, e1, e2, dump=["cost"])
    assert np.array_equal(ctl.costs_wp.numpy(), ref.dump["cost"])
    v, w = ctl.step_command("2d")                        # one-call variant, 2-D projection
    assert 0.0 <= v <= 2.0 and -1.0 <= w <= 1.0
    ctl.close()


def test_controller_run_closed_loop_reaches_a_near_goal():
    """MPPI_Controller.run (MPPI_isaac.py:755-805): the controller's own model as the plant, short hop."""
    import mppi_b200
    from mppi_b200 import MPPI_Controller, Robot, Surface
    dem, cm, hw = terrain("small")
    surface = Surface("", "", "", "", grid_size=256, half_width=12.8, origin=(0, 0), bumps=[], radius_robot=0.3)
    surface.Z = dem
    surface.costmap_size, surface.costmap_resolution, surface.costmap = 128, 25.6 / 128, np.zeros_like(cm)
    robot = Robot(x=-6.0, y=-6.0, heading_vector=[1.0, 1.0, 0.0], config_file=mppi_b200.DEFAULT_CONFIG)
    ctl = MPPI_Controller(surface, robot, mppi_b200.DEFAULT_CONFIG, goal_x=-3.0, goal_y=-3.0, goal_orientation=0.0,
                          overrides=dict(number_of_trajectories=2048, number_of_iterations=60))
    loops = ctl.run("3d", max_loops=400)
    d0 = np.hypot(-6.0 + 3.0, -6.0 + 3.0)
    d1 = np.hypot(robot.x[-1] + 3.0, robot.y[-1] + 3.0)
    assert loops > 5 and d1 < 0.75 * d0, (loops, d0, d1)          # it moves toward the goal
    assert len(robot.x) == loops + 1 and len(robot.lin_vel) == loops
    ctl.close()


def test_device_loop_replays_the_fan_of_its_last_iteration():
    """After run(device_loop=True) the replay views (.trajectories, visualiser_points, debug_dump) must show the
    rollouts the LAST executed iteration sampled -- its input state and its noise offset -- exactly as after the
    host-driven loop (round-1 advisor finding: the device loop left a stale offset and the post-step state)."""
    import mppi_b200
    from mppi_b200 import MPPI_Controller, Robot, Surface
    dem, cm, hw = terrain("small")

    def make():
        surface = Surface("", "", "", "", grid_size=256, half_width=12.8, origin=(0, 0), bumps=[], radius_robot=0.3)
        surface.Z = dem
        surface.costmap_size, surface.costmap_resolution, surface.costmap = 128, 25.6 / 128, cm
        robot = Robot(x=-6.0, y=-6.0, heading_vector=[1.0, 1.0, 0.0], config_file=mppi_b200.DEFAULT_CONFIG)
        return robot, MPPI_Controller(surface, robot, mppi_b200.DEFAULT_CONFIG, goal_x=4.0, goal_y=5.0,
                                      goal_orientation=0.0,
                                      overrides=dict(number_of_trajectories=512, number_of_iterations=40))

    rh, ch = make()
    ch.run("3d", max_loops=5)
    rd, cd = make()
    cd.run("3d", max_loops=5, device_loop=True)
    assert ch.loop == cd.loop == 5
    assert np.array_equal(np.asarray(rh.x, np.float32), np.asarray(rd.x, np.float32))
    th, td = ch.trajectories.numpy(), cd.trajectories.numpy()
    assert np.array_equal(th, td)                                     # same fan of rollouts, bit for bit
    assert np.array_equal(th[0::40][:, :2], td[0::40][:, :2])
    # ... and it IS the last iteration's fan: every rollout starts one step away from the pose BEFORE the last plant step
    start = np.array([rd.x[-2], rd.y[-2]], np.float32)
    first = td.reshape(512, 40, 3)[:, 0, :2]
    assert np.all(np.linalg.norm(first - start, axis=1) <= 2.0 * 0.045 + 1e-4)
    ph, _ = ch.visualiser_points(0.0, 0.0, 0.0)
    pd, _ = cd.visualiser_points(0.0, 0.0, 0.0)
    assert np.array_equal(ph, pd)
    ch.close()
    cd.close()


def test_two_handles_on_two_devices_in_one_process():
    """The 227 KB dynamic shared-memory opt-in is per (device, kernel): a second handle on another GPU of the same
    process must run the pipelined kernel with its 144 KB DEM tile too (round-1 advisor finding).  Needs 2 GPUs."""
    import torch
    from mppi_b200.core import Core, make_state
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    dem, cm, hw = terrain("C1")
    st = make_state(-60.57, -60.23, goal_x=65.8, goal_y=65.4)
    res = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        core = Core(4096, 100, device=d)
        core.set_terrain(torch.from_numpy(dem).to(dev), hw, torch.from_numpy(cm).to(dev))
        with torch.cuda.device(d):
            core.step(st, seed=42, offset=3, stream=torch.cuda.current_stream(dev))
            torch.cuda.synchronize(dev)
        res.append((core.optimal_u1[0].cpu().numpy().copy(), core.read_stats()))
        core.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1]["argmin"] == res[1][1]["argmin"]


def test_library_latency_ring():
    """mppi_latency_stats: p50 / p99 / max over the steps timed since mppi_enable_timing, without a profiler."""
    import torch
    core, *_ = make_core(1024, 50)
    st = state_struct(default_state())
    core.enable_timing(True)
    for i in range(12):
        core.step(st, seed=1, offset=i)
    torch.cuda.synchronize()
    s = core.latency_stats()
    assert s["n"] == 12 and 5.0 < s["p50_us"] <= s["p99_us"] <= s["max_us"] < 5000.0
    core.enable_timing(False)
    core.close()


def test_fused_peer_exchange_world1_equals_plain_step(oracle):
    """The sample-sharded step with the exchange fused into the launch (mppi_step_sharded), world = 1: the rank
    partial goes through the exchange buffer and the flag hand-shake and must reproduce mppi_step."""
    import torch
    from mppi_b200.sharding import SampleShardedStepper
    K, T = 2048, 50
    st = state_struct(default_state())
    core, *_ = make_core(K, T, lambda_=50.0)
    core.step(st, seed=3, offset=9)
    torch.cuda.synchronize()
    a = (core.optimal_u1.cpu().numpy().copy(), core.optimal_v.cpu().numpy().copy(), core.read_stats())
    stepper = SampleShardedStepper(core, K, transport="p2p")
    for _ in range(3):                                   # both buffer parities
        core.set_nominal(np.zeros(T, np.float32), np.zeros(T, np.float32))
        stepper.step(st, 3, 3, 9)
        torch.cuda.synchronize()
        b = (core.optimal_u1.cpu().numpy(), core.optimal_v.cpu().numpy(), core.read_stats())
        assert a[2]["argmin"] == b[2]["argmin"] and a[2]["min_cost"] == b[2]["min_cost"]
        assert close(b[0], a[0]) < 1e-6 and close(b[1], a[1]) < 1e-6
    core.close()


def test_fused_peer_exchange_across_gpus(tmp_path):
    """Two (or more) processes, one GPU each, launched with torchrun: p2p and NCCL transports and the unsharded
    controller agree.  Skipped on single-GPU boxes (tests/multi_gpu_check.py is the script it runs)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_CHECK_OK" in r.stdout


def test_visualiser_export_and_block_change(oracle):
    """f4 + f3: the strided visualiser feed equals the reference's slicing of the full K x T tensor
    (driver transform_trajs :252-261, :520-528), and the block-change call re-anchors history / goal and swaps maps."""
    import torch
    from mppi_b200 import DEFAULT_CONFIG, MPPI_Controller, Robot, Surface
    dem, cm, hw = terrain("small")
    surface = Surface("none", "", "none", "", dem.shape[0], hw, (0.0, 0.0), [], 0.3)
    surface.Z, surface.costmap = dem, cm
    surface.costmap_size, surface.costmap_resolution = cm.shape[0], 2 * hw / cm.shape[0]
    robot = Robot(-3.0, -2.0, (1.0, 0.2, 0.0), DEFAULT_CONFIG)
    ctrl = MPPI_Controller(surface, robot, DEFAULT_CONFIG, 6.0, 5.0, 2.2,
                           overrides=dict(number_of_trajectories=600, number_of_iterations=100))
    ctrl.warp_setup()
    ctrl.MPPI_step("3d")
    full = ctrl.trajectories.numpy()                                   # K*T x 3, as the reference keeps it
    bx, by, half = 40.0, -25.0, 87.5
    ref = full.reshape(-1, 100, 3)[::50].reshape(-1, 3)[::10]           # transform_trajs, verbatim slicing
    want = ref.copy()
    want[:, 0] = -ref[:, 1] + bx + half
    want[:, 1] = ref[:, 0] + by + half
    costs = ctrl.costs_wp.numpy()[::50]
    want_c = ((costs - np.min(costs)) / np.max(costs)).repeat(10)
    pts, cs = ctrl.visualiser_points(bx, by, half)
    assert pts.shape == want.shape and np.array_equal(pts[:, 2], want[:, 2])
    assert np.allclose(pts, want, rtol=0, atol=1e-5) and np.allclose(cs, want_c, rtol=1e-6)
    # block change: history and goal move into the new block frame, the new maps are live
    x_before, gx, gy = list(robot.x), ctrl.goal_x, ctrl.goal_y
    new_dem = torch.from_numpy(np.ascontiguousarray(dem[::-1])).cuda()
    ctrl.on_block_change(3.0, -2.0, new_dem, [(1.0, 1.0, 0.8), (-2.0, 3.0, 0.5)], (5.0, 6.0))
    assert robot.x == [x + 2.0 for x in x_before] and ctrl.goal_x == gx + 2.0 and ctrl.goal_y == gy + 3.0
    assert ctrl.Z_wp.tensor.data_ptr() == new_dem.data_ptr()             # zero-copy swap
    want_cm = oracle.obstacle_costmap([(1.0, 1.0, 0.8), (-2.0, 3.0, 0.5)], (5.0, 6.0), int(surface.costmap_size), hw, 0.3)[2]
    assert np.allclose(ctrl.surface_costmap(), want_cm, rtol=2e-5, atol=1e-7)
    ctrl.MPPI_step("3d")
    torch.cuda.synchronize()
    assert ctrl.stats()["nan"] == 0
    ctrl.close()


def test_shared_memory_dem_tile_changes_no_bit(oracle, monkeypatch):
    """The pipelined kernel stages the reachable DEM window in shared memory with a TMA tensor copy when it fits.
    With the tile switched off (MPPI_NO_DEM_TILE) the same step reads the DEM through L1 / L2: identical bits."""
    import torch
    K, T = 2048, 60
    st = state_struct(default_state())
    outs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("MPPI_NO_DEM_TILE", "1")
        else:
            monkeypatch.delenv("MPPI_NO_DEM_TILE", raising=False)
        core, *_ = make_core(K, T, lambda_=30.0, variant=2)
        core.step(st, seed=8, offset=4)
        torch.cuda.synchronize()
        outs.append((core.costs[0].cpu().numpy().copy(), core.optimal_u1[0].cpu().numpy().copy(), core.read_stats()))
        core.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2]["argmin"] == outs[1][2]["argmin"] and outs[0][2]["oob"] == 0
    # and both equal the oracle
    dem, cm, hw = terrain("C1")
    eps = oracle.philox_normals(8, 4, K, T)
    z = np.zeros(T, np.float32)
    ref = oracle.mppi_step(oracle.make_params(K=K, T=T, lam=30.0), dem, hw, cm, default_state(), z, z, eps[0], eps[1],
                           dump=["cost"])
    assert np.array_equal(outs[0][0], ref.dump["cost"])


@pytest.mark.parametrize("K,T", [(12288, 40), (300, 7), (4100, 31)])
def test_low_occupancy_instantiation_changes_no_bit(oracle, monkeypatch, K, T):
    """Single-rover launches of the monolithic kernel with at most 40960 samples run its software-pipelined
    instantiation (critics one step behind the chain, 166 registers); MPPI_NO_LOWOCC forces the throughput instantiation.
    Same costs, nominal, command and argmin, bit for bit -- and the costs are the oracle's (even / odd horizons, a K
    that is no multiple of the block)."""
    import torch
    st = state_struct(default_state())
    outs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("MPPI_NO_LOWOCC", "1")
        else:
            monkeypatch.delenv("MPPI_NO_LOWOCC", raising=False)
        core, *_ = make_core(K, T, lambda_=30.0, variant=1)
        core.step(st, seed=21, offset=6)
        torch.cuda.synchronize()
        outs.append((core.costs[0].cpu().numpy().copy(), core.optimal_u1[0].cpu().numpy().copy(),
                     core.optimal_u2[0].cpu().numpy().copy(), core.read_stats()))
        core.close()
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert np.array_equal(a, b)
    assert outs[0][3]["argmin"] == outs[1][3]["argmin"] and outs[0][3]["v0"] == outs[1][3]["v0"]
    assert outs[0][3]["w0"] == outs[1][3]["w0"]
    dem, cm, hw = terrain("C1")
    eps = oracle.philox_normals(21, 6, K, T)
    z = np.zeros(T, np.float32)
    ref = oracle.mppi_step(oracle.make_params(K=K, T=T, lam=30.0), dem, hw, cm, default_state(), z, z, eps[0], eps[1],
                           dump=["cost"])
    assert np.array_equal(outs[0][0], ref.dump["cost"])


def _dem_with_hole(ahead, lateral, rows, cols):
    dem, cm, hw = terrain("C1")
    st = default_state()
    dem = dem.copy()
    res = 2 * hw / dem.shape[0]
    i = int((st["x"] + ahead + hw) / res)                 # `ahead` metres in front of the rover (heading +x)
    j = int((hw - (st["y"] + lateral)) / res)
    dem[j:j + rows, i:i + cols] = np.nan
    return dem, cm, hw, st


@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
@pytest.mark.parametrize("math", ["strict", "fast"])
def test_nan_cells_in_the_dem_are_survivable(oracle, variant, math):
    """Real DEMs carry no-data cells.  A rollout that touches one gets NaN normals / positions; the reference would
    then gather out of bounds (it has no checks at all).  Here the step must neither fault (the shared-memory tile
    clamps its byte offset, the global path its indices) nor let the NaN into the update: such samples get zero
    weight and are counted in stats[4]; everything else stays finite.  STRICT: same NaN set and same bits as the
    oracle, over three closed-loop iterations (the controller steers away from the hole by itself)."""
    import torch
    from mppi_b200.core import Core
    K, T = 2048, 60
    dem, cm, hw, st = _dem_with_hole(2.5, 0.3, 3, 3)
    core = Core(K, T, math=math, variant=variant)
    core.set_terrain(torch.from_numpy(dem).cuda(), hw, torch.from_numpy(cm).cuda())
    n1 = n2 = np.full(T, 0.6, np.float32)
    core.set_nominal(n1, n2)
    seen = 0
    for it in range(3):
        core.step(state_struct(st), seed=3, offset=it)
        torch.cuda.synchronize()
        s = core.read_stats()
        costs = core.costs[0].cpu().numpy()
        u1 = core.optimal_u1[0].cpu().numpy()
        assert s["nan"] == int(np.isnan(costs).sum()) and s["nan"] < K
        assert np.all(np.isfinite(u1)) and np.all(np.isfinite(core.optimal_v[0].cpu().numpy()))
        assert np.isfinite(s["min_cost"]) and not np.isnan(costs[s["argmin"]])
        seen += s["nan"]
        if math == "strict":
            e1, e2 = oracle.philox_normals(3, it, K, T)
            r = oracle.mppi_step(oracle.make_params(K=K, T=T), dem, hw, cm, st, n1, n2, e1, e2, dump=["cost"],
                                 nthreads=8)
            ref = r.dump["cost"]
            assert np.array_equal(np.isnan(costs), np.isnan(ref))
            assert np.array_equal(costs[~np.isnan(ref)], ref[~np.isnan(ref)]) and s["argmin"] == r.argmin
            assert close(u1, r.nominal1, 1e-2) < RTOL
            n1, n2 = u1, core.optimal_u2[0].cpu().numpy()
    assert seen > 100
    core.close()


@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_no_valid_sample_keeps_the_nominal(variant):
    """Every rollout crosses a wall of no-data cells: no sample has a finite cost, the weights sum to zero.  The
    nominal must be kept (the reference would store 0 / 0), the command comes from it, stats[4] = K."""
    import torch
    from mppi_b200.core import Core
    K, T = 1024, 60
    dem, cm, hw, st = _dem_with_hole(1.5, 0.8, 16, 3)
    core = Core(K, T, variant=variant)
    core.set_terrain(torch.from_numpy(dem).cuda(), hw, torch.from_numpy(cm).cuda())
    n = np.full(T, 0.6, np.float32)
    core.set_nominal(n, n)
    core.step(state_struct(st), seed=3, offset=0)
    torch.cuda.synchronize()
    s = core.read_stats()
    assert s["nan"] == K and s["weights_sum"] == 0.0 and s["ess"] == 0.0
    assert np.array_equal(core.optimal_u1[0].cpu().numpy(), n) and np.array_equal(core.optimal_u2[0].cpu().numpy(), n)
    assert np.all(np.isfinite(core.optimal_v[0].cpu().numpy())) and core.optimal_v[0, 0].item() > 0
    core.close()
