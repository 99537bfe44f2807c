"""The optional critics (MppiParams.cw_orient ... cw_effort, ABI version 3).

Three of them are the reference's DORMANT critics -- `_path_orientation_critic` (critics_warp.py:44-83),
`_avoid_slope` (:131-166), `_goal_angle_critic` (:5-41): defined in the reference, their `costs[tid] +=` lines commented
out (:324, :326) or absent.  tests/golden/reference_dormant_critics.npz holds their per-sample values computed by the
reference's own functions (run under oracle/warp_shim.py by tests/golden/make_golden_critics.py) on the trajectories of
tests/golden/reference_mppi_steps.npz; the oracle and the CUDA path are held to them within the 1e-4 tolerance of the
specification.  The other three (roll, pitch, effort) are extensions without a reference counterpart: the C oracle is
cross-checked against the independent NumPy restatement, the CUDA path against the C oracle (bit-exact, STRICT).
"""
import os

import numpy as np
import pytest

from test_reference_kernels_golden import GOLD, RTOL, SCENARIOS, rel, scenario, step_io

CRIT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_dormant_critics.npz")
ALL_ON = dict(cw_orient=1.0, cw_slope_path=50.5, cw_goal_angle=30.0, cw_roll=400.0, cw_pitch=250.0, cw_effort=3.0)
GOALS = ["own", "near", "behind"]


def same_bits(a, b):
    """Bitwise equality, except that any NaN equals any NaN (payloads differ between x86 and the GPU)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def rel_nan(a, b, floor):
    a, b = np.asarray(a), np.asarray(b)
    na = np.isnan(a)
    assert np.array_equal(na, np.isnan(b))
    return rel(a[~na], b[~na], floor)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def crit():
    return np.load(CRIT)


def cases(gold, crit, name):
    sc = scenario(gold, name)
    for s in range(sc["n"]):
        for tag in GOALS:
            st, g = step_io(gold, name, s, sc)
            gx, gy = (float(v) for v in crit[f"{name}/step{s}/{tag}/goal"])
            st = dict(st, goal_x=gx, goal_y=gy, goal_theta=float(crit[f"{name}/step{s}/goal_theta"]))
            ref = dict(orient=crit[f"{name}/step{s}/{tag}/orient"], slope_path=crit[f"{name}/step{s}/slope_path"],
                       goal_angle=crit[f"{name}/step{s}/{tag}/goal_angle"])
            yield sc, s, tag, st, g, ref


def check_dormant(ext, ref, tag):
    assert rel(ext[:, 0], ref["orient"], 1e-3) < RTOL, tag
    assert rel(ext[:, 1], ref["slope_path"], 1.0) < RTOL, tag
    # a sample that stands still over the last step has atan(0 / 0) = NaN in the reference: same samples here
    nan = np.isnan(ref["goal_angle"])
    assert np.array_equal(np.isnan(ext[:, 2]), nan), tag
    assert rel(ext[~nan, 2], ref["goal_angle"][~nan], 1e-2) < RTOL, tag
    # which samples the critics fire for is decided by exact comparisons: same set as the reference's
    assert np.array_equal(ext[:, 0] != 0, ref["orient"] != 0) and np.array_equal(ext[:, 2] != 0, ref["goal_angle"] != 0)


# ------------------------------------------------------------------ CPU
def test_goldens_exercise_every_dormant_critic(crit):
    fired = {"orient": 0, "goal_angle": 0, "slope_path": 0}
    for k in crit.files:
        for n in fired:
            if k.endswith("/" + n):
                fired[n] += int(np.count_nonzero(crit[k]))
    assert all(v > 100 for v in fired.values()), fired


@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("math", ["libm", "det"])
def test_oracle_dormant_critics_match_the_reference_functions(gold, crit, oracle, name, math):
    m = oracle.MATH_LIBM if math == "libm" else oracle.MATH_DET
    for sc, s, tag, st, g, ref in cases(gold, crit, name):
        p = oracle.make_params(K=sc["K"], T=sc["T"], lam=sc["lam"], proj=sc["proj"], math=m, r_wheels=sc["radius"],
                               horizon=sc["horizon"], input_model=sc["input_model"])
        r = oracle.mppi_step(p, sc["Z"], sc["hw"], sc["cm"], st, g("in/nominal1"), g("in/nominal2"), g("out/eps1"),
                             g("out/eps2"), dump=["critics_ext", "critics", "cost"])
        check_dormant(r.dump["critics_ext"], ref, (name, s, tag))


@pytest.mark.parametrize("name", SCENARIOS)
def test_c_oracle_agrees_with_the_numpy_restatement_on_all_optional_critics(gold, crit, oracle, name):
    from oracle import mppi_oracle_np as onp
    for sc, s, tag, st, g, ref in cases(gold, crit, name):
        kw = dict(lam=sc["lam"], r_wheels=sc["radius"], horizon=sc["horizon"], **ALL_ON)
        p = oracle.make_params(K=sc["K"], T=sc["T"], proj=sc["proj"], math=oracle.MATH_LIBM,
                               input_model=sc["input_model"], **kw)
        r = oracle.mppi_step(p, sc["Z"], sc["hw"], sc["cm"], st, g("in/nominal1"), g("in/nominal2"), g("out/eps1"),
                             g("out/eps2"), dump=["critics_ext", "critics", "cost"])
        q = onp.mppi_step(onp.P(sc["K"], sc["T"], proj=sc["proj"], input_model=sc["input_model"], **kw), sc["Z"],
                          sc["hw"], sc["cm"], st, g("in/nominal1"), g("in/nominal2"), g("out/eps1"), g("out/eps2"))
        assert rel_nan(r.dump["critics_ext"], q["critics_ext"], 1e-3) < RTOL, (name, s, tag)
        assert rel_nan(r.dump["cost"], q["cost"], 1.0) < RTOL
        assert r.argmin == q["argmin"] and np.all(np.isfinite(r.nominal1)) and np.all(np.isfinite(q["nominal1"]))
        assert rel(r.nominal1, q["nominal1"], 1e-2) < RTOL
        # the documented accumulation order: orient, path, slope_path, slope, speed, obstacle, angle, roll, pitch, effort
        c4, cx = r.dump["critics"], r.dump["critics_ext"]
        f = np.float32
        c = np.zeros(sc["K"], f)
        c = c + f(ALL_ON["cw_orient"]) * cx[:, 0]
        c = c + f(100.5) * c4[:, 0]
        c = c + f(ALL_ON["cw_slope_path"]) * cx[:, 1]
        c = c + f(50.5) * c4[:, 1]
        c = c + f(0.5) * c4[:, 2]
        c = c + f(25.0) * c4[:, 3]
        for w, i in ((ALL_ON["cw_goal_angle"], 2), (ALL_ON["cw_roll"], 3), (ALL_ON["cw_pitch"], 4), (ALL_ON["cw_effort"], 5)):
            c = c + f(w) * cx[:, i]
        assert same_bits(c, r.dump["cost"])


REENABLED = dict(cw_orient=1.0, cw_slope_path=50.5)      # the coefficients of the reference's commented lines :324, :326


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_total_cost_matches_the_reference_kernel_with_its_commented_lines_reenabled(gold, crit, oracle, name):
    """tests/golden/make_golden_critics.py ran the reference's `_evaluate_trajectories_kernel` with the comment markers
    of `costs[tid] += _path_orientation_critic(...)` (:324) and `costs[tid] += 50.5*_avoid_slope(...)` (:326) removed:
    the oracle with cw_orient = 1, cw_slope_path = 50.5 reproduces those total costs, i.e. the two dormant terms sit
    where the reference would add them."""
    for sc, s, tag, st, g, ref in cases(gold, crit, name):
        want = crit[f"{name}/step{s}/{tag}/reenabled_cost"]
        p = oracle.make_params(K=sc["K"], T=sc["T"], lam=sc["lam"], proj=sc["proj"], math=oracle.MATH_LIBM,
                               r_wheels=sc["radius"], horizon=sc["horizon"], input_model=sc["input_model"], **REENABLED)
        r = oracle.mppi_step(p, sc["Z"], sc["hw"], sc["cm"], st, g("in/nominal1"), g("in/nominal2"), g("out/eps1"),
                             g("out/eps2"), dump=["cost"])
        assert rel(r.dump["cost"], want, 1.0) < RTOL, (name, s, tag)
        assert r.argmin == int(np.argmin(want))
        if tag == "own":          # and without the two weights it is the cost the unmodified kernel produced
            assert not np.array_equal(want, g("out/costs"))


def test_zero_weights_leave_the_reference_cost_untouched(gold, oracle):
    """Weight 0 means 'not evaluated, not added': an optional critic that would be NaN / inf cannot leak in."""
    name = "A3d"
    sc = scenario(gold, name)
    st, g = step_io(gold, name, 0, sc)
    st = dict(st, goal_x=float(st["x"]), goal_y=float(st["y"]))      # robot ON the goal: orient would be 0 / 0
    base = dict(K=sc["K"], T=sc["T"], lam=sc["lam"], proj=3, math=oracle.MATH_DET, r_wheels=sc["radius"],
                horizon=sc["horizon"])
    a = oracle.mppi_step(oracle.make_params(**base), sc["Z"], sc["hw"], sc["cm"], st, g("in/nominal1"),
                         g("in/nominal2"), g("out/eps1"), g("out/eps2"), dump=["cost", "critics_ext"])
    assert np.all(np.isfinite(a.dump["cost"]))
    assert not np.all(np.isfinite(a.dump["critics_ext"][:, 0])) or np.all(a.dump["critics_ext"][:, 0] == 0)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_cuda_optional_critics_bit_exact_vs_oracle_and_within_tolerance_of_the_reference(gold, crit, oracle, name,
                                                                                         variant):
    """STRICT flavour with every optional critic switched on: costs, argmin and the updated nominal equal the C oracle
    (MATH_DET) bit for bit (the nominal within 1e-4: summation order), through the fused kernels (which run the -DMPPI_XC build); the dump's per-critic values
    match the reference's own functions within 1e-4."""
    from util import GpuCore
    for sc, s, tag, st, g, ref in cases(gold, crit, name):
        kw = dict(r_wheels=sc["radius"], horizon=sc["horizon"], input_model=sc["input_model"], **ALL_ON)
        core = GpuCore(sc["K"], sc["T"], sc["Z"], sc["cm"], sc["hw"], math="strict", lambda_=sc["lam"], variant=variant,
                       **kw)
        eps = (g("out/eps1"), g("out/eps2"))
        core.set_nominal(g("in/nominal1"), g("in/nominal2"))
        res = core.step(st, proj=sc["proj"], eps=eps)
        d = core.dump(st, proj=sc["proj"], eps=eps, previous=True, names=["critics", "critics_ext"])
        p = oracle.make_params(K=sc["K"], T=sc["T"], lam=sc["lam"], proj=sc["proj"], math=oracle.MATH_DET, **kw)
        r = oracle.mppi_step(p, sc["Z"], sc["hw"], sc["cm"], st, g("in/nominal1"), g("in/nominal2"), eps[0], eps[1],
                             dump=["critics_ext", "critics", "cost"])
        key = (name, s, tag, variant)
        assert same_bits(d["critics_ext"], r.dump["critics_ext"]), key
        assert same_bits(d["critics"], r.dump["critics"]), key
        assert same_bits(res["cost"], r.dump["cost"]), key
        assert res["argmin"] == r.argmin and res["oob"] == 0
        assert res["nan"] == int(np.isnan(r.dump["cost"]).sum())       # NaN samples: zero weight, counted
        # the update sums in block order, the oracle sequentially: tolerance, as in test_gpu_parity.py
        assert rel(res["nominal1"], r.nominal1, 1e-2) < RTOL and rel(res["nominal2"], r.nominal2, 1e-2) < RTOL, key
        check_dormant(d["critics_ext"], ref, key)
        core.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENARIOS)
def test_cuda_total_cost_matches_the_reference_kernel_with_its_commented_lines_reenabled(gold, crit, name):
    """The CUDA path (STRICT, both fused kernels) against the same re-enabled reference kernel: costs within 1e-4,
    argmin exact."""
    from util import GpuCore
    for variant in (1, 2):
        for sc, s, tag, st, g, ref in cases(gold, crit, name):
            want = crit[f"{name}/step{s}/{tag}/reenabled_cost"]
            core = GpuCore(sc["K"], sc["T"], sc["Z"], sc["cm"], sc["hw"], math="strict", lambda_=sc["lam"],
                           variant=variant, r_wheels=sc["radius"], horizon=sc["horizon"],
                           input_model=sc["input_model"], **REENABLED)
            core.set_nominal(g("in/nominal1"), g("in/nominal2"))
            res = core.step(st, proj=sc["proj"], eps=(g("out/eps1"), g("out/eps2")))
            assert rel(res["cost"], want, 1.0) < RTOL, (name, s, tag, variant)
            assert res["argmin"] == int(np.argmin(want))
            core.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2], ids=["mono", "pipe"])
def test_cuda_optional_critics_fast_flavour_and_philox_noise(oracle, variant):
    """FAST flavour within tolerance of the oracle at a benchmark-like size, in-kernel Philox noise, roll / pitch /
    body-slope critics on (the set BASELINE configuration 5 names)."""
    from util import GpuCore, default_state, terrain
    K, T = 2048, 64
    dem, cm, hw = terrain("C1")
    st = default_state()
    kw = dict(cw_slope_path=50.5, cw_roll=400.0, cw_pitch=250.0)
    e1, e2 = oracle.philox_normals(42, 5, K, T)
    p = oracle.make_params(K=K, T=T, math=oracle.MATH_DET, **kw)
    z = np.zeros(T, np.float32)
    r = oracle.mppi_step(p, dem, hw, cm, st, z, z, e1, e2, dump=["cost"], nthreads=8)
    for math in ("strict", "fast"):
        core = GpuCore(K, T, dem, cm, hw, math=math, variant=variant, **kw)
        res = core.step(st, proj=3, seed=42, offset=5)
        if math == "strict":
            assert np.array_equal(res["cost"].view(np.uint32), r.dump["cost"].view(np.uint32))
            assert res["argmin"] == r.argmin
            assert rel(res["nominal1"], r.nominal1, 1e-2) < RTOL
        else:
            close = np.abs(res["cost"] - r.dump["cost"]) <= 1e-3 * np.abs(r.dump["cost"])
            assert close.mean() > 0.97
            assert abs(res["min_cost"] - r.min_cost) <= 1e-3 * abs(r.min_cost)
        core.close()


@pytest.mark.gpu
def test_facade_critic_weights_reach_the_kernels(tmp_path, oracle):
    """MPPI_Controller(critic_weights=...) changes the costs exactly as the oracle says."""
    import yaml
    from mppi_b200 import MPPI_Controller, Robot, Surface
    from util import default_state, terrain
    dem, cm, hw = terrain("C1")
    K, T = 256, 40
    cfg = dict(frame_work=dict(robot_radius=1.2), controller=dict(number_of_iterations=T, dt=0.045,
               number_of_trajectories=K), velocities=dict(initial_linear_velocity=0.0, min_linear_velocity=0.0,
               max_linear_velocity=2.0, initial_angular_velocity=0.0, min_angular_velocity=-1.0,
               max_angular_velocity=1.0), inputs=dict(std_dev_u1=0.25, std_dev_u2=0.25, min_u1=-1, max_u1=1, min_u2=-1,
               max_u2=1), cost_evaluation=dict(temperature=0.3))
    path = tmp_path / "config.yaml"
    path.write_text(yaml.safe_dump(cfg))
    st = default_state()
    surface = Surface("array", dem, "array", cm, dem.shape[0], hw, (0.0, 0.0), [], 0.3)
    costs = {}
    for tag, cw in (("base", {}), ("ext", dict(cw_pitch=250.0, cw_effort=3.0))):
        robot = Robot(st["x"], st["y"], (st["hx"], st["hy"], st["hz"]), str(path))
        c = MPPI_Controller(surface, robot, str(path), st["goal_x"], st["goal_y"], 0.0, critic_weights=cw)
        c.warp_setup()
        c.MPPI_step(proj="3d")
        costs[tag] = c.costs_wp.numpy().copy()
    assert not np.array_equal(costs["base"], costs["ext"])
    with pytest.raises(ValueError):
        MPPI_Controller(surface, robot, str(path), 0.0, 0.0, 0.0, critic_weights=dict(cw_banana=1.0))
