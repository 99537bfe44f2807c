"""GPU: the BASELINE.json configurations at their FULL sizes (C3 K = 262144, C5 K = 65536 x T = 200 on the 8192^2 DEM,
C4-shaped rover batches), checked through properties that do not need the oracle to run the whole thing:

  * oracle spot checks: the samples are keyed by GLOBAL sample id (counter-based Philox), so the CPU oracle can roll out
    any window of ids on its own -- their costs must equal the GPU's costs at those ids bit for bit (STRICT);
  * exact replay: the same (seed, offset, nominal) gives the same bits;
  * sharding invariance: K cut into ranks (step_partial + combine) gives the same costs per id and the same update;
  * the argmin the kernel reports is the argmin of the costs it wrote, sum of weights consistent with them.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def state_struct(st):
    from mppi_b200 import capi
    s = capi.MppiState()
    for k, v in st.items():
        setattr(s, k, float(v))
    return s


def build(name, **kw):
    import torch
    from mppi_b200 import synthetic as syn
    from mppi_b200.core import Core
    w = syn.WORKLOADS[name]
    dem = syn.crater_dem(w.grid_size, w.half_width, device="cuda").contiguous()
    cm = syn.rock_costmap(w.costmap_size, w.half_width)
    start, goal = syn.workload_start_goal(w)
    st = dict(x=start[0], y=start[1], hx=1.0, hy=0.0, hz=0.0, wheel_l=0.2, wheel_r=0.3, sigma1=0.25, sigma2=0.25,
              goal_x=goal[0], goal_y=goal[1], goal_theta=2.2)
    core = Core(w.K, w.T, **kw)
    core.set_terrain(dem, w.half_width, torch.from_numpy(cm).cuda())
    return w, core, dem, cm, st


def spot_check(oracle, w, dem_np, cm, st, nom, seed, offset, costs, windows, width=64, **kw):
    """Oracle rollouts of `width` consecutive sample ids starting at each window: bit-identical costs."""
    for k0 in windows:
        e1, e2 = oracle.philox_normals(seed, offset, width, w.T, k0=k0)
        p = oracle.make_params(K=width, T=w.T, math=oracle.MATH_DET, **kw)
        r = oracle.mppi_step(p, dem_np, w.half_width, cm, st, nom, nom, e1, e2, dump=["cost"], nthreads=4)
        assert np.array_equal(r.dump["cost"].view(np.uint32), costs[k0:k0 + width].view(np.uint32)), k0


def test_c3_full_size_spot_checks_replay_and_sharding(oracle):
    import torch
    w, core, dem, cm, st = build("C3")
    K, T = w.K, w.T
    s = state_struct(st)
    nom = np.full(T, 0.45, np.float32)
    core.set_nominal(nom, nom)
    core.step(s, seed=42, offset=7)
    torch.cuda.synchronize()
    costs = core.costs[0].cpu().numpy().copy()
    u1 = core.optimal_u1[0].cpu().numpy().copy()
    stats = core.read_stats()
    assert stats["nan"] == 0 and stats["oob"] == 0 and np.all(np.isfinite(costs))
    assert stats["argmin"] == int(np.argmin(costs)) and stats["min_cost"] == float(costs.min())
    wts = np.exp(-(costs.astype(np.float64) - costs.min()) / 0.3)
    assert abs(stats["weights_sum"] - wts.sum()) <= 1e-4 * wts.sum()
    spot_check(oracle, w, dem.cpu().numpy(), cm, st, nom, 42, 7, costs, [0, 4096 + 31, 131072 - 17, K - 64])
    # exact replay
    core.set_nominal(nom, nom)
    core.step(s, seed=42, offset=7)
    torch.cuda.synchronize()
    assert np.array_equal(core.costs[0].cpu().numpy(), costs) and np.array_equal(core.optimal_u1[0].cpu().numpy(), u1)
    # a different offset is a different noise stream
    core.set_nominal(nom, nom)
    core.step(s, seed=42, offset=8)
    torch.cuda.synchronize()
    assert not np.array_equal(core.costs[0].cpu().numpy()[:4096], costs[:4096])
    core.close()
    # sharding invariance: 4 ranks of K / 4 samples each (emulated on one GPU), combined in rank order
    from mppi_b200.core import Core
    G = 4
    shard = Core(K // G, T)
    shard.set_terrain(dem, w.half_width, torch.from_numpy(cm).cuda())
    parts = torch.zeros((G, shard.partial_floats()), device="cuda")
    for g in range(G):
        shard.set_nominal(nom, nom)
        shard.step_partial(s, parts[g], k_begin=g * (K // G), seed=42, offset=7)
        torch.cuda.synchronize()
        assert np.array_equal(shard.costs[0].cpu().numpy(), costs[g * (K // G):(g + 1) * (K // G)]), g
    shard.set_nominal(nom, nom)
    shard.combine_partials(s, parts, G)
    torch.cuda.synchronize()
    got = shard.read_stats()
    assert got["argmin"] == stats["argmin"] and got["min_cost"] == stats["min_cost"]
    assert np.max(np.abs(shard.optimal_u1[0].cpu().numpy() - u1)) <= 1e-6 * max(1.0, float(np.max(np.abs(u1))))
    shard.close()


def test_c5_full_size_with_slope_roll_pitch_critics(oracle):
    """BASELINE configuration 5: K = 65536, T = 200, 8192^2 DEM (268 MB), body-slope + roll + pitch critics on (the
    -DMPPI_XC kernels): oracle spot checks bit for bit, exact replay, FAST flavour within tolerance."""
    import torch
    kw = dict(cw_slope_path=50.5, cw_roll=400.0, cw_pitch=250.0)
    w, core, dem, cm, st = build("C5", **kw)
    s = state_struct(st)
    nom = np.full(w.T, 0.5, np.float32)
    core.set_nominal(nom, nom)
    core.step(s, seed=9, offset=3)
    torch.cuda.synchronize()
    costs = core.costs[0].cpu().numpy().copy()
    stats = core.read_stats()
    assert stats["nan"] == 0 and stats["oob"] == 0 and stats["argmin"] == int(np.argmin(costs))
    dem_np = dem.cpu().numpy()
    spot_check(oracle, w, dem_np, cm, st, nom, 9, 3, costs, [0, 30000 + 5, w.K - 64], **kw)
    base = oracle.make_params(K=64, T=w.T, math=oracle.MATH_DET)
    e1, e2 = oracle.philox_normals(9, 3, 64, w.T, k0=0)
    plain = oracle.mppi_step(base, dem_np, w.half_width, cm, st, nom, nom, e1, e2, dump=["cost"]).dump["cost"]
    assert not np.array_equal(plain, costs[:64])                 # the optional critics do change the cost
    core.set_nominal(nom, nom)
    core.step(s, seed=9, offset=3)
    torch.cuda.synchronize()
    assert np.array_equal(core.costs[0].cpu().numpy(), costs)
    core.close()
    del core
    from mppi_b200.core import Core
    fast = Core(w.K, w.T, math="fast", **kw)
    fast.set_terrain(dem, w.half_width, torch.from_numpy(cm).cuda())
    fast.set_nominal(nom, nom)
    fast.step(s, seed=9, offset=3)
    torch.cuda.synchronize()
    fc = fast.costs[0].cpu().numpy()
    close = np.abs(fc - costs) <= 1e-3 * np.abs(costs)
    assert close.mean() > 0.97
    fast.close()


def test_c4_shaped_rover_batch_spot_checks(oracle):
    """BASELINE configuration 4 shape (K = 1024, T = 64, 512^2 DEM + 256^2 costmap per rover), 48 rovers in one launch:
    per-rover seeds / states / maps; three rovers re-computed by the oracle bit for bit."""
    import torch
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core
    w = syn.WORKLOADS["C4"]
    R, K, T = 48, w.K, w.T
    rng = np.random.default_rng(7)
    dems = torch.stack([syn.crater_dem(w.grid_size, w.half_width, seed=100 + r, device="cuda") for r in range(R)])
    cms = np.stack([syn.rock_costmap(w.costmap_size, w.half_width, n_rocks=40, seed=200 + r) for r in range(R)])
    core = Core(K, T, max_rovers=R)
    core.set_terrain_batched(dems, w.half_width, torch.from_numpy(cms).cuda())
    half = 0.5 * w.half_width
    sts = []
    for r in range(R):
        a = rng.uniform(0, 2 * np.pi)
        sts.append(dict(x=rng.uniform(-half, half), y=rng.uniform(-half, half), hx=np.cos(a), hy=np.sin(a), hz=0.0,
                        wheel_l=0.1, wheel_r=0.2, sigma1=0.25, sigma2=0.3, goal_x=rng.uniform(-half, half),
                        goal_y=rng.uniform(-half, half), goal_theta=0.0))
    states = core.pack_states([state_struct(s) for s in sts], core.device)
    nom = np.tile(np.full(T, 0.4, np.float32), (R, 1))
    core.set_nominal(nom, nom, R)
    core.step_batched(states, R, seed=11, offset=2)
    torch.cuda.synchronize()
    costs = core.costs.cpu().numpy()
    stats = core.stats.cpu().numpy()
    for r in (0, 17, R - 1):
        f32 = {k: float(np.float32(v)) for k, v in sts[r].items()}
        e1, e2 = oracle.philox_normals(11, 2, K, T, rover=r)
        ref = oracle.mppi_step(oracle.make_params(K=K, T=T, math=oracle.MATH_DET), dems[r].cpu().numpy(), w.half_width,
                               cms[r], f32, nom[r], nom[r], e1, e2, dump=["cost"], nthreads=4)
        assert np.array_equal(ref.dump["cost"].view(np.uint32), costs[r].view(np.uint32)), r
        assert int(stats[r].view(np.int32)[1]) == ref.argmin
    core.close()
