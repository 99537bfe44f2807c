"""GPU parity tests proper: the fused sm_100a path (through the C ABI) against the CPU oracle on the same
seeded inputs.

Bars (BASELINE.json north_star):
  * bit-exact: DEM / costmap cell indices of every (k, t), and the argmin sample;
  * rollout states, costs, updated control sequence: <= 1e-4 relative (RTOL below).
In STRICT mode the kernel evaluates the same specified fp32 operation sequence as oracle/mppi_oracle.c
(MATH_DET), so states and costs are additionally required to be BIT-identical; only the softmax sums
(different, but fixed, summation order) are compared with the tolerance.
"""
import numpy as np
import pytest

from util import GpuCore, default_state, normals, terrain

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): relative error with an absolute floor for values near zero."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def run_both(oracle, K, T, name="C1", proj=3, state=None, nominal=None, seed=0, math="strict", philox=False,
             dump=True, **po):
    dem, cm, hw = terrain(name)
    st = state or default_state()
    if philox:
        eps = oracle.philox_normals(1234, 7, K, T, math=oracle.MATH_DET)
    else:
        eps = normals(K, T, seed)
    n1, n2 = nominal if nominal is not None else (np.zeros(T, np.float32), np.zeros(T, np.float32))
    okw = {("lam" if k == "lambda_" else k): v for k, v in po.items() if k != "variant"}
    p = oracle.make_params(K=K, T=T, proj=proj, math=oracle.MATH_DET, **okw)
    ref = oracle.mppi_step(p, dem, hw, cm, st, n1, n2, eps[0], eps[1], dump=True, nthreads=8)
    g = GpuCore(K, T, dem, cm, hw, math=math, **po)
    g.set_nominal(n1, n2)
    if philox:
        got = g.step(st, proj=proj, eps=None, seed=1234, offset=7)
        d = g.dump(st, proj=proj, eps=None, seed=1234, offset=7, previous=True) if dump else None
    else:
        got = g.step(st, proj=proj, eps=eps)
        d = g.dump(st, proj=proj, eps=eps, previous=True) if dump else None
    sim = g.sim_rollout(st)
    g.close()
    return ref, got, d, sim


def check_strict(ref, got, d, sim):
    assert ref.oob_clamps == 0 and got["oob"] == 0 and got["nan"] == 0
    # --- bit-exact integer work
    for n in ("dem_ij", "lw_ij", "rw_ij", "cm_ij"):
        assert np.array_equal(d[n], ref.dump[n]), f"{n} differs"
    assert got["argmin"] == ref.argmin
    # --- STRICT flavour: states and costs are bit-identical to the oracle
    for n in ("u1", "u2", "v", "w", "traj", "heading", "lw", "rw", "critics"):
        assert np.array_equal(d[n], ref.dump[n]), f"{n} not bit-identical"
    assert np.array_equal(got["cost"], ref.dump["cost"])
    assert got["min_cost"] == ref.min_cost
    assert np.array_equal(d["weights"], ref.dump["weights"])
    # --- update stage: fixed but different summation order -> tolerance
    assert rel_err(got["weights_sum"], ref.weights_sum) < RTOL
    assert rel_err(got["nominal1"], ref.nominal1) < RTOL and rel_err(got["nominal2"], ref.nominal2) < RTOL
    assert rel_err(got["nominal1"], ref.nominal1_f64) < RTOL and rel_err(got["nominal2"], ref.nominal2_f64) < RTOL
    # w* = (r - l)/track is a difference of wheel speeds: its error scales with |l|, |r| ~ |v*|, not with |w*|
    wheel_scale = max(float(np.abs(ref.opt_v).max()), 1e-3)
    assert rel_err(got["opt_v"], ref.opt_v) < RTOL and rel_err(got["opt_w"], ref.opt_w, floor=wheel_scale) < RTOL
    assert got["v0"] == got["opt_v"][0] and got["w0"] == got["opt_w"][0]
    # the optimal-trajectory rollout consumes (v*, w*): same tolerance, positions / unit headings on an absolute floor
    assert rel_err(sim[0], ref.sim_traj, floor=1.0) < RTOL and rel_err(sim[1], ref.sim_heading, floor=1.0) < RTOL


MONO, PIPE = 1, 2          # MPPI_VARIANT_*: both fused kernels must produce the same bits
both_variants = pytest.mark.parametrize("variant", [MONO, PIPE], ids=["mono", "pipe"])


@both_variants
@pytest.mark.parametrize("K,T", [(1024, 50), (4096, 100), (1000, 100), (37, 7), (1, 2), (33, 9)])
def test_strict_3d_injected_noise(oracle, K, T, variant):
    """C1 / C2 shapes (+ ragged K, tiny and odd T) with shared injected noise."""
    ref, got, d, sim = run_both(oracle, K, T, "C1", variant=variant)
    check_strict(ref, got, d, sim)


@both_variants
def test_strict_2d_injected_noise(oracle, variant):
    ref, got, d, sim = run_both(oracle, 1024, 50, "C1", proj=2, variant=variant)
    check_strict(ref, got, d, sim)
    # the 2-D kernel leaves the wheel arrays at zero: slope critic = number of terms
    assert np.all(d["critics"][:, 1] == 24.0)


@both_variants
def test_strict_philox_production_mode(oracle, variant):
    """Production mode: in-kernel Philox4x32-10 + det Box-Muller equals the oracle's restatement of the stream."""
    ref, got, d, sim = run_both(oracle, 2048, 100, "C1", philox=True, variant=variant)
    check_strict(ref, got, d, sim)


@both_variants
def test_rough_terrain_with_lethal_cells(oracle, variant):
    """Small bumpy map with dense rocks: exercises lethal costmap cells, steep wheel slopes, near-goal branch off."""
    st = default_state(x=-5.0, y=-4.0, hx=0.6, hy=0.8, hz=0.0, goal_x=8.0, goal_y=9.0, wheel_l=0.4, wheel_r=0.7)
    n1 = np.linspace(0.9, 0.2, 64).astype(np.float32)
    n2 = np.linspace(0.5, 0.8, 64).astype(np.float32)
    ref, got, d, sim = run_both(oracle, 2048, 64, "small", state=st, nominal=(n1, n2), seed=3, variant=variant)
    check_strict(ref, got, d, sim)
    assert (ref.dump["critics"][:, 3] > 1e5).any(), "scenario should hit lethal cells"


@both_variants
def test_noisy_dem_takes_the_general_normalisation_path(oracle, variant):
    """2 cm of cell-to-cell noise on the DEM: the slope under the rover changes by degrees between steps, so the
    tangent projection leaves the +-2^-15 window of the MUFU-free shortcut on (nearly) every step, lanes diverge
    between the two code paths, and the wheel slopes are large.  Still bit-identical to the oracle."""
    n = np.full(80, 0.7, np.float32)
    ref, got, d, sim = run_both(oracle, 2048, 80, "rough", nominal=(n, n), seed=11, variant=variant, philox=True)
    check_strict(ref, got, d, sim)
    # the scenario does what it says: |h . n| of consecutive steps is far outside the shortcut's window
    hd, tr = ref.dump["heading"], ref.dump["traj"]
    dz = np.abs(np.diff(hd[:, :, 2], axis=1))
    assert np.median(dz) > 0.02 and np.max(np.abs(tr[:, -1, 0] - tr[:, 0, 0])) > 1.0


@both_variants
def test_near_goal_branches(oracle, variant):
    """dist < horizon -> path-follow near branch (sum of L1 distances); dist < 2 -> speed critic off."""
    n = np.full(50, 0.5, np.float32)
    st = default_state(x=-2.0, y=1.0, goal_x=0.5, goal_y=2.5)            # 2.9 m: near branch, speed on
    ref, got, d, sim = run_both(oracle, 512, 50, "small", state=st, nominal=(n, n), seed=4, variant=variant)
    check_strict(ref, got, d, sim)
    assert np.all(ref.dump["critics"][:, 2] != 0.0)
    st = default_state(x=-2.0, y=1.0, goal_x=-1.0, goal_y=1.5)           # 1.1 m: speed critic returns 0
    ref, got, d, sim = run_both(oracle, 512, 50, "small", state=st, nominal=(n, n), seed=5, variant=variant)
    check_strict(ref, got, d, sim)
    assert np.all(ref.dump["critics"][:, 2] == 0.0)


@both_variants
def test_high_temperature_exercises_weighted_update(oracle, variant):
    """Large lambda -> effective sample size >> 1 so the weighted sum (not just argmin) is tested."""
    n = np.full(50, 0.3, np.float32)
    ref, got, d, sim = run_both(oracle, 4096, 50, "C1", nominal=(n, n), seed=6, lambda_=3.0e4, variant=variant)
    check_strict(ref, got, d, sim)
    assert got["ess"] > 100.0
    w = ref.dump["weights"].astype(np.float64)
    assert abs(got["ess"] - w.sum() ** 2 / (w ** 2).sum()) / got["ess"] < 1e-3


def test_near_tie_argmin(oracle):
    """Duplicate the best sample's noise into a later slot: exact tie -> lowest index must win (np.argmin)."""
    K, T = 1024, 50
    dem, cm, hw = terrain("C1")
    st = default_state()
    e1, e2 = normals(K, T, 11)
    z = np.zeros(T, np.float32)
    p = oracle.make_params(K=K, T=T)
    r0 = oracle.mppi_step(p, dem, hw, cm, st, z, z, e1, e2, dump=["cost"])
    best = r0.argmin
    dup = (best + 517) % K
    e1[dup], e2[dup] = e1[best], e2[best]
    ref = oracle.mppi_step(p, dem, hw, cm, st, z, z, e1, e2, dump=True)
    g = GpuCore(K, T, dem, cm, hw)
    got = g.step(st, eps=(e1, e2))
    g.close()
    assert ref.argmin == min(best, dup) == got["argmin"]
    assert rel_err(got["nominal1"], ref.nominal1) < RTOL


def test_fast_flavour_within_tolerance(oracle):
    """FAST arithmetic (FMA contraction, approximate div/rsqrt): tolerance on states and costs, no bit claims."""
    K, T = 2048, 100
    ref, got, d, sim = run_both(oracle, K, T, "C1", math="fast")
    # a flipped cell index is possible when a point sits within an ulp of a cell border: allow a handful
    same = np.all(d["dem_ij"] == ref.dump["dem_ij"], axis=-1).mean()
    assert same > 0.999
    assert rel_err(d["traj"][..., :2], ref.dump["traj"][..., :2]) < RTOL
    ok = np.isclose(got["cost"], ref.dump["cost"], rtol=1e-3, atol=1e-2)
    assert ok.mean() > 0.995          # a flipped costmap/DEM cell changes that sample's cost legitimately


def test_closed_loop_sequence_stays_bit_identical(oracle):
    """Ten consecutive iterations, each consuming the previous nominal (receding horizon) and the simulated pose:
    GPU and oracle run side by side; costs stay bit-identical as long as the nominal matches to rounding."""
    K, T = 1024, 50
    dem, cm, hw = terrain("C1")
    g = GpuCore(K, T, dem, cm, hw)
    st = default_state()
    p = oracle.make_params(K=K, T=T)
    n1 = np.zeros(T, np.float32)
    n2 = np.zeros(T, np.float32)
    for it in range(10):
        eps = oracle.philox_normals(99, it, K, T)
        g.set_nominal(n1, n2)                       # keep both sides on the same nominal: isolates one step
        got = g.step(st, eps=None, seed=99, offset=it)
        ref = oracle.mppi_step(p, dem, hw, cm, st, n1, n2, eps[0], eps[1], dump=["cost"])
        assert np.array_equal(got["cost"], ref.dump["cost"]) and got["argmin"] == ref.argmin
        assert rel_err(got["nominal1"], ref.nominal1) < RTOL
        n1, n2 = ref.nominal1, ref.nominal2
        st = default_state(x=float(ref.sim_traj[0, 0]), y=float(ref.sim_traj[0, 1]),
                           hx=float(ref.sim_heading[0, 0]), hy=float(ref.sim_heading[0, 1]),
                           hz=float(ref.sim_heading[0, 2]),
                           wheel_l=float(ref.opt_v[0] - ref.opt_w[0] * 0.6),
                           wheel_r=float(ref.opt_v[0] + ref.opt_w[0] * 0.6))
    g.close()


@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_rollouts_leaving_the_map_are_clamped_like_the_oracle(oracle, variant):
    """Start 0.6 m from the map edge, heading out, wheels already spinning: most samples leave the DEM.  The
    reference has no bounds checks (undefined behaviour); the core clamps and counts -- the kernels take their
    clamped code path here (terrain_window_safe is false) and must still reproduce the oracle's clamped costs bit for
    bit, cell indices included, and report the violations."""
    K, T = 512, 40
    dem, cm, hw = terrain("small")
    st = default_state(x=hw - 0.6, y=-hw + 0.9, hx=1.0, hy=-0.3, goal_x=0.0, goal_y=0.0, wheel_l=1.8, wheel_r=1.8)
    nom = (np.full(T, 0.9, np.float32), np.full(T, 0.9, np.float32))
    ref, got, d, sim = run_both(oracle, K, T, "small", state=st, nominal=nom, variant=variant)
    assert ref.oob_clamps > 0 and got["oob"] > 0 and got["nan"] == 0
    for n in ("dem_ij", "lw_ij", "rw_ij", "cm_ij"):
        assert np.array_equal(d[n], ref.dump[n]), f"{n} differs"
    assert np.array_equal(got["cost"], ref.dump["cost"]) and got["argmin"] == ref.argmin
    assert rel_err(got["nominal1"], ref.nominal1_f64) < RTOL
