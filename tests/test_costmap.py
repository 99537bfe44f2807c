"""f1 -- obstacle-costmap builder (Surface.create_obstacles_costmap, MPPI_isaac.py:361-378).

CPU: the oracle restatement (oracle/costmap_oracle.c) against the reference recipe evaluated with the reference's own
dependencies (NumPy meshgrid mask, cv2.distanceTransform / cv2.normalize -- both are in this image, on the GPU box too):
identical mask; distance map within one float32 ulp of cv2 (cv2 4.13 dispatches DIST_L2 / 5 to Intel IPP, whose
rounding of exact ties differs from the plain two-pass float recursion in a few cells per ten thousand).
GPU: mppi_build_costmap against the oracle (mask and distance map BIT-identical) and against cv2.
"""
import numpy as np
import pytest


def reference_recipe(obstacles, origin, cms, hw, r_robot, power=20):
    """The reference's lines MPPI_isaac.py:361-378 with its own libraries (restated calls, not copied kernels)."""
    import cv2
    xc = np.linspace(-hw, hw, cms)
    X, Y = np.meshgrid(xc, xc)
    obs = 255 * np.ones((cms, cms), dtype=np.uint8)
    for xg, yg, r in obstacles:
        xl, yl = yg - origin[1], xg - origin[0]
        R = r / 2 + r_robot + 0.1
        obs[(X - xl) ** 2 + (Y - yl) ** 2 <= R ** 2] = 0
    d = cv2.distanceTransform(obs, cv2.DIST_L2, 5)
    dn = cv2.normalize(d, None, 0, 1.0, cv2.NORM_MINMAX)
    return obs, d, (1 - dn) ** power


def rocks(n, span, seed):
    rng = np.random.default_rng(seed)
    return [(float(rng.uniform(-span, span)), float(rng.uniform(-span, span)), float(rng.uniform(0.1, 1.6)))
            for _ in range(n)]


def ulp_close(a, b, ulps=1):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)))))


CASES = [(200, 20.0, 40, 3, (1.5, -2.0)), (93, 9.3, 12, 5, (0.0, 0.0)), (320, 32.0, 150, 8, (-3.0, 4.0))]


@pytest.mark.parametrize("cms,hw,n,seed,origin", CASES)
def test_oracle_matches_the_reference_recipe(oracle, cms, hw, n, seed, origin):
    obst = rocks(n, 0.7 * hw, seed)
    mask_r, d_r, c_r = reference_recipe(obst, origin, cms, hw, 0.3)
    mask_o, d_o, c_o = oracle.obstacle_costmap(obst, origin, cms, hw, 0.3)
    assert np.array_equal(mask_o, mask_r)
    assert ulp_close(d_o, d_r, 1) and (d_o != d_r).mean() < 2e-3
    assert np.allclose(c_o, c_r, rtol=2e-5, atol=1e-7)


def test_chamfer_restatement_against_cv2_on_random_masks(oracle):
    import cv2
    rng = np.random.default_rng(0)
    for n, p in [(64, 0.02), (200, 0.001), (33, 0.2), (400, 0.0005)]:
        m = (rng.random((n, n)) > p).astype(np.uint8) * 255
        a, b = cv2.distanceTransform(m, cv2.DIST_L2, 5), oracle.chamfer5x5(m)
        assert ulp_close(a, b, 1) and (a != b).mean() < 2e-3
    m = np.full((40, 60), 255, np.uint8)
    m[5, 7] = 0
    assert np.array_equal(cv2.distanceTransform(m, cv2.DIST_L2, 5), oracle.chamfer5x5(m))


@pytest.mark.gpu
@pytest.mark.parametrize("cms,hw,n,seed,origin", CASES + [(875, 87.5, 750, 11, (2.0, -1.0)), (750, 75.0, 0, 1, (0.0, 0.0))])
def test_gpu_builder_matches_oracle_and_cv2(oracle, cms, hw, n, seed, origin):
    import torch
    from mppi_b200 import build_obstacle_costmap
    obst = rocks(n, 0.7 * hw, seed)
    c_g, d_g, m_g = build_obstacle_costmap(obst, origin, cms, hw, 0.3, want_intermediates=True)
    torch.cuda.synchronize()
    c_g, d_g, m_g = c_g.cpu().numpy(), d_g.cpu().numpy(), m_g.cpu().numpy()
    mask_o, d_o, c_o = oracle.obstacle_costmap(obst, origin, cms, hw, 0.3)
    assert np.array_equal(m_g, mask_o)                       # identical rasterisation
    assert np.array_equal(d_g, d_o)                          # bit-identical chamfer distances
    assert np.allclose(c_g, c_o, rtol=2e-5, atol=1e-7)       # pow: double-rounded-once vs NumPy's float32 powf
    if n:
        mask_r, d_r, c_r = reference_recipe(obst, origin, cms, hw, 0.3)
        assert np.array_equal(m_g, mask_r)
        assert ulp_close(d_g, d_r, 1)
        assert np.allclose(c_g, c_r, rtol=2e-5, atol=1e-7)


@pytest.mark.gpu
def test_controller_rebuild_costmap_feeds_the_next_step(oracle, tmp_path):
    """Driver block-change path: rebuild on the device, then step -- the step must see the new costmap."""
    import torch
    from mppi_b200 import DEFAULT_CONFIG, MPPI_Controller, Robot, Surface
    from util import terrain
    dem, _, hw = terrain("small")
    surface = Surface("none", "", "none", "", dem.shape[0], hw, (0.0, 0.0), [], 0.3)
    surface.Z = dem
    robot = Robot(-3.0, -2.0, (1.0, 0.2, 0.0), DEFAULT_CONFIG)
    ctrl = MPPI_Controller(surface, robot, DEFAULT_CONFIG, 6.0, 5.0, 2.2,
                           overrides=dict(number_of_trajectories=512, number_of_iterations=40))
    ctrl.warp_setup()
    ctrl.MPPI_step("3d")
    torch.cuda.synchronize()
    c0 = ctrl.costs_wp.numpy().copy()
    obst = rocks(30, 8.0, 3) + [(-2.0, -2.6, 1.5)]            # a rock right in front of the rover
    ctrl.rebuild_costmap(obst, (0.0, 0.0))
    ref = oracle.obstacle_costmap(obst, (0.0, 0.0), int(surface.costmap_size), hw, 0.3)[2]
    assert np.allclose(ctrl.surface_costmap(), ref, rtol=2e-5, atol=1e-7)
    ctrl._step_count = 0
    ctrl.optimal_u1_wp.tensor.zero_(); ctrl.optimal_u2_wp.tensor.zero_()
    ctrl.MPPI_step("3d")
    torch.cuda.synchronize()
    assert ctrl.costs_wp.numpy().mean() > c0.mean()           # obstacles cost something now
    ctrl.close()
