"""CPU: pins the oracle against golden vectors produced by the REFERENCE's own CPU code
(tests/golden/make_golden.py executed thesis_master/python_mppi_projection/displacement_on_surface.py) and
against the one golden file the reference ships (trajectory_2D.csv)."""
import os

import numpy as np
import pytest

from oracle import mppi_oracle_np as onp

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_functions.npz"))
f32 = np.float32


def test_normal_on_grid_matches_reference():
    q = G["q"].astype(f32)
    n = onp.normal_on_grid(q[:, 0, 0], q[:, 0, 1], q[:, 1, 0], q[:, 1, 1], f32(G["res"]))
    np.testing.assert_allclose(n, G["normals"], rtol=1e-5, atol=1e-6)


def test_tangent_matches_reference():
    t = onp.tangent(G["normals"].astype(f32), G["heads"].astype(f32))
    np.testing.assert_allclose(t, G["tangents"], rtol=1e-5, atol=2e-6)


def test_position_and_rodrigues_match_reference_rotvec():
    """Reference CPU update_position = position step + SciPy rotvec about the normal; the Warp kernels spell
    the rotation out as Rodrigues' formula (projection_warp.py:240-244)."""
    tg, n = G["tangents"].astype(f32), G["normals"].astype(f32)
    x, y = onp.update_position(G["xy"][:, 0].astype(f32), G["xy"][:, 1].astype(f32), tg, G["v"].astype(f32),
                               f32(G["dt"]))
    h = onp.update_orientation(tg, G["w"].astype(f32), n, f32(G["dt"]))
    np.testing.assert_allclose(x, G["upd"][:, 0], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(y, G["upd"][:, 1], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(h, G["upd"][:, 2:5], rtol=1e-5, atol=2e-6)


def test_bilinear_matches_reference_for_positive_coordinates():
    q = G["q"].astype(f32)
    b = onp.bilinear(G["xy_pos"][:, 0].astype(f32), G["xy_pos"][:, 1].astype(f32), q[:, 0, 0], q[:, 0, 1], q[:, 1, 0],
                     q[:, 1, 1], f32(G["res"]))
    # x/res is evaluated in fp32 at |x/res| ~ 150: the fraction carries ~1e-5 absolute error
    np.testing.assert_allclose(b, G["bil"], rtol=0, atol=2e-4)


def test_bilinear_negative_coordinate_quirk():
    """trunc-based fractions go negative for negative coordinates: the Warp semantics differ from
    the CPU script's floor there -- the oracle follows Warp."""
    q = [f32(0.0), f32(1.0), f32(2.0), f32(3.0)]        # q00 q01 q10 q11
    pos = onp.bilinear(f32(0.03), f32(0.0), *q, f32(0.1))
    neg = onp.bilinear(f32(-0.03), f32(0.0), *q, f32(0.1))
    assert abs(pos - 0.6) < 1e-5 and abs(neg + 0.6) < 1e-5          # x-fraction weights the ROW neighbour q10


def _rollout_2d(oracle, v, w, x0, y0, h, dt):
    T = len(v)
    p = onp.P(1, T, proj=2, dt=dt)
    ter = onp.Terrain(np.zeros((64, 64), f32), 1000.0, np.zeros((8, 8), f32))
    return onp.rollout(p, ter, f32(x0), f32(y0), np.asarray(h, f32), np.asarray(v, f32)[None], np.asarray(w, f32)[None])


def test_2d_rollout_matches_reference_generate_trajectory_2D(oracle):
    ref = G["traj2"]                                   # ref[0] is the start; ref[k+1] follows step k
    ro = _rollout_2d(oracle, G["v2"], G["w2"], -3.0, 2.0, G["h2"], 0.045)
    ours = ro["traj"][0]
    np.testing.assert_allclose(ours[:-1, :2], ref[1:, :2], rtol=0, atol=5e-5)


def test_reference_trajectory_2D_csv():
    """The only golden file in the reference: v = 1.5, w = 0, dt = 0.01 from (-14, -4) -> x_k = -14 + 0.015 k."""
    n = int(G["csv_len"])
    ro = _rollout_2d(None, np.full(n, 1.5), np.zeros(n), -14.0, -4.0, [1.0, 0.0, 0.0], 0.01)
    ours = ro["traj"][0]
    idx = G["csv_idx"]
    rows = G["csv_rows"]
    sel = idx[idx >= 1]
    # csv row k = position after k steps; ours[t] = position after t+1 steps.  fp32 accumulation over 3000 steps.
    np.testing.assert_allclose(ours[sel - 1, 0], rows[idx >= 1, 0], rtol=0, atol=2e-3)
    np.testing.assert_allclose(ours[sel - 1, 1], rows[idx >= 1, 1], rtol=0, atol=1e-6)
    assert rows[0, 0] == -14.0 and rows[0, 1] == -4.0
