"""CPU: host-side mirror of the reference interface (Surface / Robot / MPPI_Controller construction, config
parsing, state packing) and the partitioning helpers."""
import numpy as np
import pytest

import mppi_b200
from mppi_b200 import MPPI_Controller, Robot, Surface, capi, synthetic
from mppi_b200.sharding import shard_range


def test_surface_geometry_follows_the_reference():
    s = Surface("manual", "", "manual", "", grid_size=400, half_width=20.0, origin=(0.0, 0.0),
                bumps=[((1.0, 2.0), 2.0, 3.0)], radius_robot=0.3, obstacles=[(3.0, 4.0, 0.5)])
    assert s.resolution == pytest.approx(0.1) and s.costmap_size == 50 and s.costmap_resolution == pytest.approx(0.8)
    assert s.Z.shape == (400, 400) and s.costmap.shape == (50, 50)
    # crater formula MPPI_isaac.py:317-320 at the crater centre: (h - 0.5) - (h + 0.5) = -1
    x = np.linspace(-20, 20, 400)
    i, j = np.argmin(np.abs(x - 1.0)), np.argmin(np.abs(x - 2.0))
    assert s.Z[j, i] == pytest.approx(-1.0, abs=2e-2)
    assert 0.0 <= s.costmap.min() and s.costmap.max() == pytest.approx(1.0)


def test_robot_and_controller_read_the_yaml_schema():
    r = Robot(1.0, 2.0, [3.0, 4.0, 0.0], mppi_b200.DEFAULT_CONFIG)
    assert r.radius == 1.2 and np.allclose(r.heading_vector, [0.6, 0.8, 0.0])
    r.update_position(1.5, 2.5, 0.1, np.array([0.0, 1.0, 0.0]))
    assert r.x[-1] == 1.5 and r.z[-1] == 0.1
    s = Surface("", "", "", "", grid_size=64, half_width=3.2, origin=(0, 0), bumps=[], radius_robot=0.3)
    c = MPPI_Controller(s, r, mppi_b200.DEFAULT_CONFIG, goal_x=5.0, goal_y=6.0, goal_orientation=2.2)
    assert (c.number_of_trajectories, c.number_of_iterations, c.dt) == (1000, 100, 0.045)
    assert c.horizon == pytest.approx(9.0) and c.temperature == 0.3
    p = c._params()
    assert (p.K, p.T) == (1000, 100) and p.r_wheels == pytest.approx(1.2) and p.horizon == pytest.approx(9.0)
    st = c._state()
    assert (st.x, st.y, st.hx, st.hy) == pytest.approx((1.5, 2.5, 0.0, 1.0))
    assert (st.goal_x, st.goal_y, st.sigma1) == pytest.approx((5.0, 6.0, 0.25))
    with pytest.raises(capi.MppiError):
        c.MPPI_step("3d")                      # warp_setup() not called
    c.reset("controller")                      # no-op, kept for drop-in compatibility


def test_reference_config_file_is_accepted_if_present():
    import os
    ref = "/root/reference/thesis_master/warp_implementation/config.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    r = Robot(0.0, 0.0, [1.0, 0.0, 0.0], ref)
    s = Surface("", "", "", "", grid_size=64, half_width=3.2, origin=(0, 0), bumps=[], radius_robot=0.3)
    c = MPPI_Controller(s, r, ref, 1.0, 1.0, 0.0)
    ours = MPPI_Controller(s, r, mppi_b200.DEFAULT_CONFIG, 1.0, 1.0, 0.0)
    for k in ("number_of_iterations", "dt", "number_of_trajectories", "std_dev_u1", "std_dev_u2", "temperature",
              "v_max_linear", "v_min_angular", "min_u1", "max_u2"):
        assert getattr(c, k) == getattr(ours, k), k


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 4096, 262144, 262147):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(total, world, r) for r in range(world)]
            assert sum(n for _, n in pieces) == total
            pos = 0
            for b, n in pieces:
                assert b == pos
                pos += n
            assert max(n for _, n in pieces) - min(n for _, n in pieces) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_synthetic_workloads_have_the_baseline_shapes():
    w = synthetic.WORKLOADS
    assert (w["C1"].K, w["C1"].T) == (1024, 50) and (w["C2"].K, w["C2"].T) == (4096, 100)
    assert (w["C2"].grid_size, w["C2"].costmap_size) == (1500, 750)
    assert (w["C3"].K, w["C3"].grid_size) == (262144, 2048) and w["C4"].n_rovers == 4096
    assert (w["C5"].K, w["C5"].T, w["C5"].grid_size) == (65536, 200, 8192)
    cm = synthetic.rock_costmap(64, 6.4, n_rocks=5, seed=1)
    assert cm.shape == (64, 64) and cm.dtype == np.float32 and cm.max() == 1.0
    dem = synthetic.crater_dem(64, 6.4, bumps=[((0.0, 0.0), 1.0, 1.0)])
    assert tuple(dem.shape) == (64, 64)


def test_bench_reference_arm_emits_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path on the host cores) runs without a GPU and prints
    ONE JSON line with the keys the measurement contract names."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MPPI sample-steps/s" and d["unit"] == "sample-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C1" in d["config"]["workload"]


def test_surface_reproduces_the_reference_surface_arrays():
    """tests/golden/reference_mppi_steps.npz stores the DEM and the obstacle costmap the REFERENCE's `Surface`
    (MPPI_isaac.py:259-378, run under the Warp shim by make_golden_warp.py) built for the golden scenarios; the facade's
    `Surface` with the same arguments gives the same DEM bit for bit and the same costmap (cv2 on both sides)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_mppi_steps.npz"))
    bumps = [((-1.5, 1.0), 1.4, 2.5), ((3.0, -1.0), 2.0, 3.0), ((0.5, 4.0), 0.9, 1.5), ((-4.0, -4.0), 1.2, 2.0)]
    obstacles = [(1.0, 2.0, 0.6), (-2.0, -1.0, 0.8), (3.5, 3.0, 0.5), (0.0, -3.0, 0.7), (-3.0, 3.5, 0.4)]
    s = Surface("manual", "", "manual", "", 160, 8.0, (0.0, 0.0), bumps, 0.3, obstacles)
    assert np.array_equal(np.asarray(s.Z, np.float32), gold["A3d/Z"])
    assert np.allclose(np.asarray(s.costmap, np.float32), gold["A3d/costmap"], rtol=2e-5, atol=1e-7)
    assert (s.grid_size, s.costmap_size) == (160, int(gold["A3d/meta"][4]))
    assert s.resolution == pytest.approx(float(gold["A3d/fmeta"][1])) and \
        s.costmap_resolution == pytest.approx(float(gold["A3d/fmeta"][2]))


def _reference_module():
    """thesis_master/warp_implementation/MPPI_isaac.py imported under the Warp shim (no GPU needed to construct its
    classes), or a skip when the reference tree is not mounted."""
    import os
    import sys
    import types
    ref_root = "/root/reference"
    cfg = os.path.join(ref_root, "thesis_master/warp_implementation/config.yaml")
    if not os.path.exists(cfg):
        pytest.skip("reference tree not mounted")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import warp_shim
    saved = {k: sys.modules.get(k) for k in ("warp", "matplotlib", "matplotlib.pyplot")}
    try:
        sys.modules["warp"] = warp_shim
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        sys.path.insert(0, ref_root)
        import thesis_master.warp_implementation.MPPI_isaac as ref
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if ref_root in sys.path:
            sys.path.remove(ref_root)
    return ref, cfg


def test_facade_classes_agree_with_the_reference_classes_if_present():
    """Live comparison (build container only): the reference's own `Surface`, `Robot` and `MPPI_Controller`
    (MPPI_isaac.py, imported under the Warp shim -- constructing them needs no GPU) against the facade's, attribute
    by attribute, after the same constructor calls and pose updates."""
    ref, cfg = _reference_module()
    bumps = [((-1.0, 0.5), 1.2, 2.0), ((2.0, -1.5), 1.8, 2.5)]
    rocks = [(1.0, 1.0, 0.5), (-2.0, 0.5, 0.7)]
    args = ("manual", "", "manual", "", 96, 4.8, (0.0, 0.0), bumps, 0.3, rocks)
    a, b = ref.Surface(*args), Surface(*args)
    for k in ("grid_size", "half_width", "resolution", "costmap_size", "costmap_resolution"):
        assert getattr(a, k) == getattr(b, k), k
    assert np.array_equal(np.asarray(a.Z, np.float32), np.asarray(b.Z, np.float32))
    assert np.allclose(a.costmap, b.costmap, rtol=2e-5, atol=1e-7)
    ra, rb = ref.Robot(0.3, -0.4, [3.0, 4.0, 0.0], cfg), Robot(0.3, -0.4, [3.0, 4.0, 0.0], cfg)
    for r in (ra, rb):
        r.update_position(0.5, -0.2, 0.1, np.array([0.0, 1.0, 0.0]))
    for k in ("x", "y", "z", "radius", "left_wheel_speed", "right_wheel_speed", "lin_vel", "ang_vel"):
        assert getattr(ra, k) == getattr(rb, k), k
    assert np.array_equal(ra.heading_vector, rb.heading_vector)
    ca, cb = ref.MPPI_Controller(a, ra, cfg, 3.0, 2.0, 2.2), MPPI_Controller(b, rb, cfg, 3.0, 2.0, 2.2)
    for k in ("goal_x", "goal_y", "goal_orientation", "loop", "number_of_iterations", "dt", "number_of_trajectories",
              "initial_linear_velocity", "initial_angular_velocity", "std_dev_u1", "std_dev_u2", "min_u1", "max_u1",
              "min_u2", "max_u2", "v_min_linear", "v_max_linear", "v_min_angular", "v_max_angular", "temperature",
              "horizon"):
        assert getattr(ca, k) == getattr(cb, k), k


def test_visualiser_transform_matches_the_drivers_transform_trajs_if_present():
    """The driver's `transform_trajs` (visual_terrain_stack_full_terrain.py:252-261) cannot be imported (the file pulls
    in Isaac Sim), but the function itself is plain NumPy: its source is cut out with `ast` and executed.  Its output
    on a K x 100 x 3 array equals the facade's world-frame transform applied to the every-50th-sample /
    every-10th-step subset, i.e. to what `mppi_export_trajectories` produces (that equality is a GPU test)."""
    import ast
    import os
    import torch
    path = "/root/reference/visual_terrain_stack_full_terrain.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    src = open(path).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "transform_trajs")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)

    class FakeWarpArray:
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    rng = np.random.default_rng(3)
    K, T = 1000, 100
    traj = rng.uniform(-20, 20, (K * T, 3)).astype(np.float32)
    bx, by, hb = 123.25, -77.5, 20.0
    want = ns["transform_trajs"](FakeWarpArray(traj), bx, by, hb, 0.0)
    sub = torch.from_numpy(traj.reshape(K, T, 3)[::50, ::10].reshape(-1, 3).copy())
    got = MPPI_Controller.to_world_frame(sub, bx, by, hb).numpy()
    assert got.shape == want.shape == (20 * 10, 3) and np.array_equal(got, want)
    c = torch.from_numpy(rng.uniform(100, 5000, K).astype(np.float32))[::50]
    cn = c.numpy()
    assert np.array_equal(MPPI_Controller.normalised_costs(c).numpy(), (cn - np.min(cn)) / np.max(cn))


def test_synthetic_costmap_recipe_is_the_reference_recipe_if_present():
    """bench.py's obstacle costmaps come from `synthetic.costmap_from_free_mask`, which claims to be the reference's
    offline recipe.  That recipe exists only as commented lines in create_costmap.py:14-28: here they are un-commented
    in memory (np.load / np.save redirected to arrays) and executed on the same binary map -- identical costmap."""
    import os
    path = "/root/reference/thesis_master/warp_implementation/create_costmap.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    lines = open(path).read().split("\n")[13:28]
    code = [l[2:] if l.startswith("# ") else l for l in lines if l.strip() and not l.startswith("# #")]
    assert any("cv2.distanceTransform" in l for l in code) and any("**10" in l for l in code)
    free = synthetic.rock_free_mask(256, 25.6, n_rocks=40, seed=5)
    binary = (free == 0).astype(np.uint8)                       # the reference's map: 1 = obstacle, 0 = free space
    saved = {}

    class FakeNp:
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def load(name):
            return binary

        @staticmethod
        def save(name, a):
            saved[name] = a

    import cv2
    exec("\n".join(code), {"np": FakeNp(), "cv2": cv2})
    assert list(saved) == ["costmap_750_transformed.npy"]
    ours = synthetic.costmap_from_free_mask(free, 10.0)
    assert np.array_equal(ours, saved["costmap_750_transformed.npy"].astype(np.float32))
    assert np.array_equal(ours, synthetic.rock_costmap(256, 25.6, n_rocks=40, seed=5))


def test_synthetic_scene_constants_are_the_references_if_present():
    """The bench scene quotes the reference's own experiment constants: the nine craters, the start and
    goal, and the rock generator (RandomState(99), 750 rocks within +-50 m of a 75 m map, r in U(0, 0.4), disc radius
    r + r_robot + 0.2) are read back from the text of MPPI_OO_current.py and compared."""
    import ast
    import os
    import re
    path = "/root/reference/thesis_master/warp_implementation/MPPI_OO_current.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    text = open(path).read()
    m = re.search(r"# bumps = \[\n((?:#\s+\(\(.*\n)+)# \]", text)
    assert m
    craters = ast.literal_eval("[" + "".join(l.lstrip("# ") for l in m.group(1).splitlines()) + "]")
    assert craters == synthetic.NINE_CRATERS
    nums = {k: float(re.search(r"#\s+#\s+%s = (-?[0-9.]+)" % k, text).group(1))
            for k in ("x_start", "y_start", "x_goal", "y_goal")}
    w = synthetic.WORKLOADS["C2"]
    assert (nums["x_start"], nums["y_start"]) == w.start and (nums["x_goal"], nums["y_goal"]) == w.goal
    assert "rng = np.random.RandomState(99)" in text and "for i in range(750):" in text
    assert "rng.uniform(-50.0, 50.0), rng.uniform(-50.0, 50.0), rng.uniform(0.0, 0.4)" in text
    assert "(r_obs + self.r_robot + 0.2) ** 2" in text
    # the generator draws in the same order from the same stream: first rock of the 75 m map
    rng = np.random.RandomState(99)
    first = (rng.uniform(-50.0, 50.0), rng.uniform(-50.0, 50.0), rng.uniform(0.0, 0.4))
    free = synthetic.rock_free_mask(750, 75.0, n_rocks=1)
    xc = np.linspace(-75.0, 75.0, 750)
    want = ((xc[None, :] - first[0]) ** 2 + (xc[:, None] - first[1]) ** 2) <= (first[2] + 0.3 + 0.2) ** 2
    assert np.array_equal(free == 0, want)


def test_synthetic_crater_dem_is_the_reference_surface_if_present():
    """The C1 / C2 bench DEM (`synthetic.crater_dem(1500, 75.0)`: the nine craters) against the reference's own
    `Surface.create_surface` (MPPI_isaac.py:300-322) for the same arguments.  The generator here adds each crater only
    within +-6 widths of its centre (beyond that the Gaussians are < 2e-8 of the depth), hence 1e-6 instead of bits."""
    ref, _ = _reference_module()
    theirs = ref.Surface("manual", "", "", "", 1500, 75.0, (0.0, 0.0), synthetic.NINE_CRATERS, 0.3)
    ours = synthetic.crater_dem(1500, 75.0).numpy()
    assert ours.shape == np.asarray(theirs.Z).shape == (1500, 1500)
    assert np.max(np.abs(ours - np.asarray(theirs.Z, np.float32))) < 1e-6
    assert np.ptp(ours) > 4.0                                   # a real crater field (rims ~3.9 m, floors ~-1 m), not a plane


def _fracs(obj, path=""):
    if isinstance(obj, dict):
        for k, v in obj.items():
            if k == "frac" and isinstance(v, (int, float)):
                yield path, v
            else:
                yield from _fracs(v, f"{path}/{k}")


def test_committed_bench_lines_carry_the_measurement_contract():
    """The lines measured on B200s this round (profiles/r2_bench/) have every key the measurement contract names, with
    consistent values -- a schema check of what bench.py's GPU arm emits (that arm cannot run here).  No roofline
    fraction may exceed 1 (round 1 printed issue.frac = 2.78 from counters of another configuration), the bound is named
    after what binds, and the same-run sub-objects carry passed equality checks."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = {}
    for path in sorted(glob.glob(os.path.join(root, "profiles", "r2_bench", "*.json"))):
        lines[os.path.basename(path)] = json.loads([l for l in open(path) if l.startswith("{")][0])
    assert "c2_strict_default.json" in lines and len(lines) >= 10
    for name, d in lines.items():
        if d.get("impl") == "reference":
            assert d["gpu_launches"] == 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
            continue
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert k in d, (name, k)
        assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["warmup"] >= 3
        assert "model" not in d["config"] and "workload" in d["config"]
        r = d["roofline"]
        assert r["bound"] in ("latency", "issue") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        for where, f in _fracs(r):
            assert 0.0 <= f <= 1.0, (name, where, f)
        if "issue" in r:
            assert r["issue"]["extrapolated"] in (False, True)
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["e2e"]["value"] < d["value"] * 1.001
    d = lines["c2_strict_default.json"]
    assert d["n_gpus"] == 1 and "C2" in d["config"]["workload"] and d["gpu_launches"] == d["steps"]
    assert abs(d["value"] - 4096 * 100 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["roofline"]["bound"] == "latency" and d["roofline"]["traffic"] > 0 and d["roofline"]["issue"]["extrapolated"] is False
    assert 0.5 < d["roofline"]["kernel_share_of_step"] <= 1.0                    # measured, not asserted
    assert d["e2e"]["h2d_bytes_per_step"] == 48 and d["e2e"]["d2h_bytes_per_step"] == 8
    assert d["latency_us"]["steps"] >= 1000                                     # p50 / p99 do not depend on --steps
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["reference_script"]["kind"] == "reference"
    assert cb["reference_script"]["cores"] == 1 and cb["reference_script"]["value"] < cb["value"]
    assert d["ms_per_step"] * 1e3 < 100.0                                       # BASELINE target: C2 under 100 us
    sc = d["sharded_check"]
    assert sc["passed"] and sc["p2p_bitwise"] and sc["argmin_equal"] and sc["lambdas"] == [0.3, 2000.0]
    assert d["c3"]["weak"]["check"]["passed"] and d["c3"]["weak"]["K_per_gpu"] == 262144
    assert d["c4"]["check"]["passed"] and d["c4"]["rovers_per_gpu"] == 512
    # the multi-GPU lines: same per-GPU workload, more GPUs
    for n in (2, 4, 8):
        m = lines[f"c2_strict_n{n}.json"]
        assert m["n_gpus"] == n and m["config"]["K_total"] == n * 4096 and m["scaling"] == "weak"
        assert m["value"] > lines["c2_strict_n1.json"]["value"]
    assert lines["c5many_strict_ext.json"]["roofline"]["hbm_measured"]["frac"] > 0.05   # the many-start C5 leaves L2
