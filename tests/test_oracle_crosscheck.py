"""CPU: the two independent restatements of the reference step (C scalar, NumPy vectorised) must agree, and
both must satisfy closed-form known answers (flat DEM, inclined plane, quirks)."""
import numpy as np
import pytest

from oracle import mppi_oracle_np as onp
from util import default_state, normals, terrain

f32 = np.float32


def both(oracle, K, T, name="C1", proj=3, state=None, nominal=None, seed=0, math=None, **kw):
    dem, cm, hw = terrain(name) if isinstance(name, str) else name
    st = state or default_state()
    e1, e2 = normals(K, T, seed)
    n1, n2 = nominal if nominal is not None else (np.zeros(T, f32), np.zeros(T, f32))
    p = oracle.make_params(K=K, T=T, proj=proj, math=oracle.MATH_LIBM if math is None else math, **kw)
    c = oracle.mppi_step(p, dem, hw, cm, st, n1, n2, e1, e2, dump=True, nthreads=4)
    n = onp.mppi_step(onp.P(K, T, proj=proj, **kw), dem, hw, cm, st, n1, n2, e1, e2)
    return c, n


@pytest.mark.parametrize("proj", [3, 2])
def test_c_and_numpy_restatements_agree(oracle, proj):
    K, T = 256, 50
    nom = (np.full(T, 0.4, f32), np.full(T, 0.6, f32))
    c, n = both(oracle, K, T, "C1", proj=proj, nominal=nom, seed=1)
    for name in ("u1", "u2", "v", "w"):
        assert np.array_equal(c.dump[name], n[name]), name            # no transcendentals: bit-identical
    # sin/cos differ by <= 1 ulp between glibc and NumPy: positions agree to rounding
    np.testing.assert_allclose(c.dump["traj"], n["traj"], rtol=1e-6, atol=2e-5)
    np.testing.assert_allclose(c.dump["heading"], n["heading"], rtol=0, atol=2e-6)
    same_idx = np.all(c.dump["dem_ij"] == n["dem_ij"], axis=-1).mean()
    assert same_idx > 0.999
    agree = np.isclose(c.dump["cost"], n["cost"], rtol=1e-4, atol=1e-2)
    assert agree.mean() > 0.99
    assert c.argmin == n["argmin"]
    np.testing.assert_allclose(c.nominal1, n["nominal1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(c.opt_v, n["opt_v"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(c.sim_traj, n["sim_traj"], rtol=1e-5, atol=2e-5)


def test_det_and_libm_math_agree_to_tolerance(oracle):
    """The specified transcendental functions change results only at rounding level (north_star tolerance 1e-4)."""
    K, T = 512, 100
    dem, cm, hw = terrain("C1")
    st = default_state()
    e1, e2 = normals(K, T, 2)
    nom = np.full(T, 0.5, f32)
    res = {}
    for m in (oracle.MATH_LIBM, oracle.MATH_DET):
        res[m] = oracle.mppi_step(oracle.make_params(K=K, T=T, math=m), dem, hw, cm, st, nom, nom, e1, e2, dump=True)
    a, b = res[oracle.MATH_LIBM], res[oracle.MATH_DET]
    np.testing.assert_allclose(a.dump["traj"], b.dump["traj"], rtol=1e-6, atol=2e-5)
    assert np.all(a.dump["dem_ij"] == b.dump["dem_ij"], axis=-1).mean() > 0.999
    assert a.argmin == b.argmin
    np.testing.assert_allclose(a.nominal1, b.nominal1, rtol=1e-4, atol=1e-6)


def flat_maps(height=0.7, hw=20.0, gs=400):
    return np.full((gs, gs), height, f32), np.zeros((gs // 8, gs // 8), f32), hw


def test_flat_dem_known_answers(oracle):
    """Z = const: n = (0,0,1), the 3-D rollout equals the 2-D one and the closed-form unicycle arc; height = const;
    wheel slopes are 0 so the slope critic is the number of stride-2 pairs (T even: (T-2)/2)."""
    K, T = 64, 40
    maps = flat_maps()
    st = default_state(x=-3.0, y=1.0, hx=0.6, hy=0.8, goal_x=15.0, goal_y=12.0)
    nom = (np.full(T, 0.5, f32), np.full(T, 0.7, f32))
    c3, n3 = both(oracle, K, T, maps, proj=3, state=st, nominal=nom, seed=3)
    c2, n2 = both(oracle, K, T, maps, proj=2, state=st, nominal=nom, seed=3)
    np.testing.assert_allclose(c3.dump["traj"][..., :2], c2.dump["traj"][..., :2], rtol=0, atol=1e-5)
    # the four bilinear weights sum to 1 only up to fp32 rounding
    assert np.abs(c3.dump["traj"][..., 2] - 0.7).max() < 2e-7 and np.abs(n3["traj"][..., 2] - 0.7).max() < 2e-7
    assert np.all(c3.dump["critics"][:, 1] == (T - 2) / 2) and np.all(c2.dump["critics"][:, 1] == (T - 2) / 2)
    # closed-form: theta_{t+1} = theta_t + w_t dt ; p_{t+1} = p_t + v_t dt (cos theta_t, sin theta_t)
    v, w = c3.dump["v"].astype(np.float64), c3.dump["w"].astype(np.float64)
    th = np.arctan2(0.8, 0.6) + np.concatenate([np.zeros((K, 1)), np.cumsum(w * 0.045, axis=1)[:, :-1]], axis=1)
    x = -3.0 + np.cumsum(v * 0.045 * np.cos(th), axis=1)
    y = 1.0 + np.cumsum(v * 0.045 * np.sin(th), axis=1)
    np.testing.assert_allclose(c3.dump["traj"][..., 0], x, rtol=0, atol=2e-5)
    np.testing.assert_allclose(c3.dump["traj"][..., 1], y, rtol=0, atol=2e-5)
    assert np.abs(np.linalg.norm(c3.dump["heading"], axis=-1) - 1).max() < 1e-6


def test_inclined_plane_known_answers(oracle):
    """Z varies linearly along the grid: the quad normal is constant, headings stay unit and tangent to it."""
    gs, hw = 400, 20.0
    res = 2 * hw / gs
    jj, ii = np.meshgrid(np.arange(gs), np.arange(gs), indexing="ij")
    dem = (0.02 * ii - 0.01 * jj).astype(f32)          # +0.02 per column (x), -0.01 per row (row index grows with -y)
    maps = (dem, np.zeros((50, 50), f32), hw)
    st = default_state(x=-2.0, y=-1.0, hx=1.0, hy=0.0, goal_x=15.0, goal_y=12.0)
    nom = (np.full(30, 0.6, f32), np.full(30, 0.5, f32))
    c, n = both(oracle, 32, 30, maps, proj=3, state=st, nominal=nom, seed=4, math=oracle.MATH_DET)
    # closed form of projection_warp.py:142-151 for this plane: (q01-q00) = 0.02, (q10-q00) = -0.01
    vec = np.array([-res / 2 * 0.04, -res / 2 * -0.02, res * res])
    nrm = vec / np.linalg.norm(vec)
    hd = c.dump["heading"].astype(np.float64)
    assert np.abs(hd @ nrm).max() < 2e-5                      # tangent to the plane
    assert np.abs(np.linalg.norm(hd, axis=-1) - 1).max() < 1e-6


def test_row_flip_and_index_formula(oracle):
    """projection_warp.py:39-40: i = int((x + hw)/res), j = -int((y - hw)/res): row 0 is y = +hw."""
    gs, hw = 200, 10.0
    dem = np.zeros((gs, gs), f32)
    ter = onp.Terrain(dem, hw, np.zeros((25, 25), f32))
    i, j = onp.cell_index(ter, np.array([-10.0, 0.0, 9.99, 3.05], f32), np.array([10.0, 0.0, -9.95, -2.51], f32))
    assert list(i) == [0, 100, 199, 130] and list(j) == [0, 100, 199, 125]
    # C oracle reports the same indices through a rollout with v = 0 (stays at the start)
    st = default_state(x=3.05, y=-2.51, goal_x=9.0, goal_y=9.0)
    p = oracle.make_params(K=1, T=4)
    z = np.zeros((1, 4), f32)
    r = oracle.mppi_step(p, dem, hw, np.zeros((25, 25), f32), st, np.full(4, -1, f32), np.full(4, -1, f32), z, z,
                         dump=True)
    assert np.all(r.dump["dem_ij"][0] == [130, 125])
    assert np.all(r.dump["cm_ij"][0] == [int((3.05 + 10) / 0.8), int((2.51 + 10) / 0.8)])


def test_update_stage_given_costs(oracle):
    """Softmax update in isolation: weights / nominal against float64, including the argmin-collapse regime."""
    K, T = 300, 20
    c, n = both(oracle, K, T, "small", state=default_state(x=-5, y=-4, goal_x=8, goal_y=9), seed=5, lam=5000.0)
    cost = c.dump["cost"].astype(np.float64)
    w = np.exp(-(cost - cost.min()) / 5000.0)
    nom = (w[:, None] * c.dump["u1"].astype(np.float64)).sum(0) / w.sum()
    np.testing.assert_allclose(c.nominal1, nom, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(c.nominal1_f64, nom, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(n["nominal1"], nom, rtol=1e-4, atol=1e-6)
    assert (w > 1e-3).sum() > 10
    # reference settings: lambda = 0.3 collapses onto the argmin sample
    c2, _ = both(oracle, K, T, "small", state=default_state(x=-5, y=-4, goal_x=8, goal_y=9), seed=5)
    np.testing.assert_allclose(c2.nominal1, c2.dump["u1"][c2.argmin], rtol=1e-5, atol=1e-6)


def test_combine_partials_is_invariant_to_the_split(oracle):
    """Online-softmax partials of any contiguous split fold back to the unsharded update."""
    K, T = 512, 30
    c, _ = both(oracle, K, T, "small", state=default_state(x=-5, y=-4, goal_x=8, goal_y=9), seed=6, lam=800.0,
                math=oracle.MATH_DET)
    cost, u1, u2 = c.dump["cost"], c.dump["u1"], c.dump["u2"]
    for G in (1, 2, 4, 8):
        parts = np.zeros((G, 4 + 2 * T), f32)
        for g in range(G):
            sl = slice(g * K // G, (g + 1) * K // G)
            m = cost[sl].min()
            w = np.exp(-(cost[sl].astype(np.float64) - m) / 800.0)
            parts[g, 0], parts[g, 1], parts[g, 3] = m, w.sum(), (w ** 2).sum()
            parts[g, 2] = np.array([sl.start + int(np.argmin(cost[sl]))], np.int32).view(f32)[0]
            parts[g, 4:4 + T] = (w[:, None] * u1[sl]).sum(0)
            parts[g, 4 + T:] = (w[:, None] * u2[sl]).sum(0)
        # the C combine takes {M, S, argmin, A1, A2} = stride 3 + 2T
        packed = np.concatenate([parts[:, :3], parts[:, 4:]], axis=1)
        n1, n2, m, arg, s = oracle.combine_partials(packed, T, 800.0)
        assert arg == c.argmin and m == c.min_cost
        np.testing.assert_allclose(n1, c.nominal1_f64, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(n2, c.nominal2_f64, rtol=2e-6, atol=1e-7)


def test_velocity_space_input_model_c_vs_numpy(oracle):
    """input_model = 1 (unicycle: _generate_velocities_kernel, sampling_warp.py:10-48): (v, w) are sampled directly
    around the previous optimal velocity sequence and clamped to the velocity limits; no wheel filter anywhere."""
    K, T = 192, 40
    nom = (np.full(T, 0.8, f32), np.linspace(-0.3, 0.3, T).astype(f32))
    st = default_state(sigma1=0.3, sigma2=0.2, wheel_l=0.7, wheel_r=0.1)       # wheel speeds must be ignored
    c, n = both(oracle, K, T, "C1", nominal=nom, seed=4, state=st, input_model=1, lam=500.0)
    e1, e2 = normals(K, T, 4)
    src = np.minimum(np.arange(T) + 1, T - 1)
    v_expect = np.clip(nom[0][src][None, :] + f32(0.3) * e1, f32(0.0), f32(2.0)).astype(f32)
    assert np.array_equal(c.dump["v"], v_expect) and np.array_equal(c.dump["u1"], c.dump["v"])
    assert np.array_equal(c.dump["w"], n["w"]) and np.array_equal(c.dump["u2"], c.dump["w"])
    assert c.argmin == n["argmin"]
    np.testing.assert_allclose(c.dump["cost"], n["cost"], rtol=2e-5)
    np.testing.assert_allclose(c.nominal1, n["nominal1"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(c.opt_v, c.nominal1) and np.array_equal(c.opt_w, c.nominal2)


def test_oracle_ctypes_mirrors_follow_the_oracle_header():
    """oracle/oracle_c.py mirrors the structures of oracle/mppi_oracle.h by hand: field names, order and types are
    compared here, so that a new field cannot silently shift the checker's arguments."""
    import ctypes as C
    import os
    import re
    from oracle import oracle_c
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "oracle", "mppi_oracle.h")).read(), flags=re.S)
    ctype = {"int32_t": C.c_int32, "float": C.c_float, "double": C.c_double}
    rename = {"lambda": "lam"}
    for name in ("OrParams", "OrTerrain", "OrState", "OrDump", "OrOut"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(?:const )?(\w+) (.*)", decl)
            for item in m.group(2).split(","):
                item = item.strip()
                n = item.lstrip("*").strip()
                fields.append((rename.get(n, n), C.c_void_p if item.startswith("*") else ctype[m.group(1)]))
        assert list(getattr(oracle_c, name)._fields_) == fields, name


def test_oracle_result_does_not_depend_on_its_thread_count(oracle):
    """The CPU baseline runs the oracle on all host cores (bench.py); its outputs must be the same bits as the
    single-threaded run the parity tests use (static partition of the samples, sequential reductions)."""
    from util import default_state, normals, terrain
    dem, cm, hw = terrain("small")
    K, T = 333, 40
    st = default_state(x=-5.0, y=-4.0, hx=0.6, hy=0.8, goal_x=8.0, goal_y=9.0)
    e1, e2 = normals(K, T, 9)
    n = np.full(T, 0.4, np.float32)
    p = oracle.make_params(K=K, T=T, lam=40.0, cw_pitch=250.0)
    runs = [oracle.mppi_step(p, dem, hw, cm, st, n, n, e1, e2, dump=["cost", "traj", "critics_ext"], nthreads=t)
            for t in (1, 3, 16)]
    for r in runs[1:]:
        assert np.array_equal(r.dump["cost"], runs[0].dump["cost"]) and np.array_equal(r.dump["traj"], runs[0].dump["traj"])
        assert np.array_equal(r.dump["critics_ext"], runs[0].dump["critics_ext"])
        assert np.array_equal(r.nominal1, runs[0].nominal1) and np.array_equal(r.opt_w, runs[0].opt_w)
        assert (r.argmin, r.min_cost, r.weights_sum, r.oob_clamps) == (runs[0].argmin, runs[0].min_cost,
                                                                      runs[0].weights_sum, runs[0].oob_clamps)
