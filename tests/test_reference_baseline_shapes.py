"""ONE step of the reference's own `MPPI_Controller.MPPI_step` (MPPI_isaac.py:505-720 + sampling_warp.py,
projection_warp.py, critics_warp.py, imported unmodified, interpreted by oracle/warp_shim.py; generator:
tests/golden/make_golden_warp_shapes.py) at the shapes BASELINE.json names, against

  * the C oracle (CPU tests): pins the restatement to the reference's kernels AT C1 / C2 SIZE, not only at K <= 64;
  * the CUDA path through the C ABI (`-m gpu`): both fused kernels, STRICT -- shared injected noise, bit-exact
    u / v / omega and argmin, <= 1e-4 relative on states, costs, the updated control sequence and the command.

Scenarios: C1 / C2 (bench start, flat ground outside the rock field), C1rock / C2rock (inside the rock field on a
crater wall, warm nominal, unequal sigmas: lethal cells and real slopes).  The fixture stores the terrain WINDOW the
rollouts can reach; it is pasted into NaN-filled full-size maps, so any read outside the window would poison the
result.  The noise is recomputed from the stored seed with the shim's generator and checked against a stored checksum.
"""
import functools
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_mppi_baseline_shapes.npz")
SCENARIOS = ["C1", "C1rock", "C2", "C2rock"]
RTOL = 1e-4          # north_star: rollout states, costs and the updated control sequence within 1e-4 relative


@functools.lru_cache(maxsize=None)
def load(name):
    from oracle import warp_shim as wp
    gold = np.load(GOLD)
    g = lambda k: gold[f"{name}/{k}"]                                            # noqa: E731
    K, T, gs, cms, sub, seed, di0, di1, dj0, dj1, ci0, ci1, cj0, cj1 = (int(x) for x in g("meta"))
    hw, res, cres, lam, gx, gy, horizon, radius, sum1, sum2 = (float(x) for x in g("fmeta"))
    Z = np.full((gs, gs), np.nan, np.float32)
    Z[dj0:dj1, di0:di1] = g("Z_window")
    cm = np.full((cms, cms), np.nan, np.float32)
    cm[cj0:cj1, ci0:ci1] = g("costmap_window")
    # the noise the reference's sampling kernel drew from the shim's wp.randn (state formula: sampling_warp.py:71-92)
    tid = np.arange(K * T, dtype=np.int64)
    last = (tid % T) == (T - 1)
    s1 = np.where(last, seed + tid + 3 * T, seed + tid + T)
    s2 = np.where(last, seed + tid + 4 * T, seed + tid + 2 * T)
    e1 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s1], np.float32).reshape(K, T)
    e2 = np.array([wp.randn_from_state(wp.uint32(s)) for s in s2], np.float32).reshape(K, T)
    assert float(e1.astype(np.float64).sum()) == sum1 and float(e2.astype(np.float64).sum()) == sum2
    h = g("in/heading")
    st = dict(x=g("in/x"), y=g("in/y"), hx=h[0], hy=h[1], hz=h[2], wheel_l=g("in/wheel_l"), wheel_r=g("in/wheel_r"),
              sigma1=g("in/sigma1"), sigma2=g("in/sigma2"), goal_x=gx, goal_y=gy, goal_theta=2.2)
    return dict(K=K, T=T, sub=sub, hw=hw, lam=lam, horizon=horizon, radius=radius, Z=Z, cm=cm, eps=(e1, e2), st=st, g=g)


def rel(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def check(sc, got):
    """got: full K x T arrays u1 u2 v w traj heading lw rw + cost argmin weights_sum nominal1/2 opt_v/w."""
    g, sub = sc["g"], slice(0, sc["K"], sc["sub"])
    for k in ("u1", "u2", "v", "w"):                       # no transcendental function before (v, omega): same bits
        assert np.array_equal(got[k][sub], g("out/" + k)), k
    assert rel(got["traj"][sub], g("out/traj"), 1e-2) < RTOL
    assert rel(got["heading"][sub], g("out/heading_vectors"), 1e-2) < RTOL
    assert rel(got["lw"][sub], g("out/lw"), 1e-2) < RTOL and rel(got["rw"][sub], g("out/rw"), 1e-2) < RTOL
    ref_cost = g("out/costs")
    assert not np.isnan(got["cost"]).any()                 # nothing was read outside the stored terrain window
    assert rel(got["cost"], ref_cost, 1.0) < RTOL
    assert got["argmin"] == int(np.argmin(ref_cost))       # bit-exact argmin sample
    assert rel(got["weights_sum"], g("out/weights_sum"), 1e-3) < RTOL
    assert rel(got["nominal1"], g("out/out_nominal1"), 1e-2) < RTOL
    assert rel(got["nominal2"], g("out/out_nominal2"), 1e-2) < RTOL
    assert rel(got["opt_v"], g("out/opt_v"), 1e-2) < RTOL and rel(got["opt_w"], g("out/opt_w"), 1e-2) < RTOL


def test_fixture_covers_what_it_claims():
    """The rock-field scenarios really cross lethal cells and slopes; the bench start does not."""
    for name in SCENARIOS:
        sc = load(name)
        traj, cm, hw = sc["g"]("out/traj"), sc["cm"], sc["hw"]
        cres = 2 * hw / cm.shape[0]
        ix = ((traj[..., 0] + hw) / cres).astype(np.int64)                      # critics_warp.py:245-248
        iy = ((-traj[..., 1] + hw) / cres).astype(np.int64)
        lethal = (cm[iy, ix] > 0.99).any(axis=1)
        dz = np.ptp(traj[..., 2])
        if name.endswith("rock"):
            assert 0 < lethal.sum() < lethal.size and dz > 0.15
        else:
            assert lethal.sum() == 0 and dz < 0.05


@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("math", ["libm", "det"])
def test_oracle_matches_the_reference_step_at_baseline_shapes(oracle, name, math):
    sc = load(name)
    g = sc["g"]
    p = oracle.make_params(K=sc["K"], T=sc["T"], lam=sc["lam"], proj=3, r_wheels=sc["radius"], horizon=sc["horizon"],
                           math=oracle.MATH_LIBM if math == "libm" else oracle.MATH_DET)
    r = oracle.mppi_step(p, sc["Z"], sc["hw"], sc["cm"], sc["st"], g("in/nominal1"), g("in/nominal2"), *sc["eps"],
                         dump=True)
    got = dict(r.dump, argmin=r.argmin, weights_sum=r.weights_sum, nominal1=r.nominal1, nominal2=r.nominal2,
               opt_v=r.opt_v, opt_w=r.opt_w)
    check(sc, got)
    assert rel(r.dump["weights"], g("out/weights"), 1e-3) < RTOL
    assert r.min_cost == pytest.approx(float(g("out/min_cost")), rel=RTOL)
    assert rel(r.sim_traj, g("out/sim_traj"), 1e-2) < RTOL and rel(r.sim_heading, g("out/sim_heading"), 1e-2) < RTOL
    assert r.oob_clamps == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_cuda_path_matches_the_reference_step_at_baseline_shapes(name, variant):
    from util import GpuCore
    sc = load(name)
    g = sc["g"]
    core = GpuCore(sc["K"], sc["T"], sc["Z"], sc["cm"], sc["hw"], math="strict", lambda_=sc["lam"],
                   r_wheels=sc["radius"], horizon=sc["horizon"], variant=variant)
    core.set_nominal(g("in/nominal1"), g("in/nominal2"))
    res = core.step(sc["st"], proj=3, eps=sc["eps"])
    d = core.dump(sc["st"], proj=3, eps=sc["eps"], previous=True,
                  names=["u1", "u2", "v", "w", "traj", "heading", "lw", "rw"])
    got = dict(d, cost=res["cost"], argmin=res["argmin"], weights_sum=res["weights_sum"], nominal1=res["nominal1"],
               nominal2=res["nominal2"], opt_v=res["opt_v"], opt_w=res["opt_w"])
    check(sc, got)
    assert res["oob"] == 0 and res["nan"] == 0
    sim_t, sim_h = core.sim_rollout(sc["st"])
    assert rel(sim_t, g("out/sim_traj"), 1e-2) < RTOL and rel(sim_h, g("out/sim_heading"), 1e-2) < RTOL
    core.close()
