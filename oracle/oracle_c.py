"""TEST INFRASTRUCTURE -- ctypes binding of the C oracle (oracle/mppi_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmppi_oracle.so")

MATH_LIBM = 0
MATH_DET = 1


class OrParams(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("T", C.c_int32), ("proj", C.c_int32), ("math", C.c_int32),
        ("dt", C.c_float),
        ("u1_min", C.c_float), ("u1_max", C.c_float), ("u2_min", C.c_float), ("u2_max", C.c_float),
        ("v_min", C.c_float), ("v_max", C.c_float), ("w_min", C.c_float), ("w_max", C.c_float),
        ("lam", C.c_float), ("r_wheels", C.c_float),
        ("filt_k", C.c_float), ("filt_a", C.c_float), ("opt_k", C.c_float), ("opt_a", C.c_float),
        ("wheel_offset", C.c_float),
        ("cw_path", C.c_float), ("cw_slope", C.c_float), ("cw_speed", C.c_float), ("cw_obs", C.c_float),
        ("lethal_thresh", C.c_float), ("lethal_penalty", C.c_float),
        ("near_goal_cut", C.c_float), ("speed_eps", C.c_float), ("pf_eps", C.c_float),
        ("pf_near_gain", C.c_float), ("slope_eps", C.c_float), ("slope_gain", C.c_float),
        ("horizon", C.c_float), ("target_speed", C.c_float), ("input_model", C.c_int32),
        ("cw_orient", C.c_float), ("cw_slope_path", C.c_float), ("cw_goal_angle", C.c_float),
        ("goal_angle_radius", C.c_float), ("cw_roll", C.c_float), ("cw_pitch", C.c_float), ("cw_effort", C.c_float),
    ]


class OrTerrain(C.Structure):
    _fields_ = [
        ("dem", C.c_void_p), ("gs", C.c_int32), ("half_width", C.c_float), ("res", C.c_float),
        ("costmap", C.c_void_p), ("cms", C.c_int32), ("cres", C.c_float),
    ]


class OrState(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("x", "y", "hx", "hy", "hz", "wheel_l", "wheel_r", "sigma1", "sigma2",
                 "goal_x", "goal_y", "goal_theta")]


class OrDump(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("u1", "u2", "v", "w", "traj", "heading", "lw", "rw",
                 "dem_ij", "lw_ij", "rw_ij", "cm_ij", "critics", "critics_ext", "cost", "weights")]


class OrOut(C.Structure):
    _fields_ = [
        ("nominal1", C.c_void_p), ("nominal2", C.c_void_p), ("opt_v", C.c_void_p), ("opt_w", C.c_void_p),
        ("sim_traj", C.c_void_p), ("sim_heading", C.c_void_p),
        ("nominal1_f64", C.c_void_p), ("nominal2_f64", C.c_void_p),
        ("min_cost", C.c_float), ("argmin", C.c_int32), ("weights_sum", C.c_float), ("oob_clamps", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    srcs = [os.path.join(_HERE, f) for f in ("mppi_oracle.c", "mppi_oracle.h", "det_math.h", "costmap_oracle.c", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libmppi_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_int32]
        L.oracle_philox_normals.restype = None
        L.oracle_detmath_eval.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.oracle_detmath_eval.restype = None
        L.oracle_mppi_step.argtypes = [C.POINTER(OrParams), C.POINTER(OrTerrain), C.POINTER(OrState),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(OrDump), C.POINTER(OrOut), C.c_int32]
        L.oracle_mppi_step.restype = C.c_int
        L.oracle_combine_partials.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_combine_partials.restype = None
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def philox_normals(seed: int, offset: int, K: int, T: int, rover: int = 0, k0: int = 0, math: int = MATH_DET):
    e1 = np.zeros((K, T), dtype=np.float32)
    e2 = np.zeros((K, T), dtype=np.float32)
    lib().oracle_philox_normals(seed, offset, rover, k0, K, T, _p(e1), _p(e2), math)
    return e1, e2


def detmath(fn: int, x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y0 = np.zeros_like(x)
    y1 = np.zeros_like(x)
    lib().oracle_detmath_eval(fn, _p(x), _p(y0), _p(y1), x.size)
    return y0, y1


# Reference defaults (config.yaml and literals in the Warp kernels).
DEFAULTS = dict(
    K=1000, T=100, proj=3, math=MATH_DET, dt=0.045,
    u1_min=-1.0, u1_max=1.0, u2_min=-1.0, u2_max=1.0,
    v_min=0.0, v_max=2.0, w_min=-1.0, w_max=1.0,
    lam=0.3, r_wheels=1.2, filt_k=3.5, filt_a=0.96, opt_k=3.0, opt_a=0.92, wheel_offset=0.2,
    cw_path=100.5, cw_slope=50.5, cw_speed=0.5, cw_obs=25.0,
    lethal_thresh=0.99, lethal_penalty=100000.0,
    near_goal_cut=2.0, speed_eps=0.0001, pf_eps=1e-6, pf_near_gain=10.0, slope_eps=1e-6, slope_gain=5.0,
    goal_angle_radius=0.5,                # critics_warp.py:33; the optional critics' weights default to 0 (off)
)


def make_params(**kw) -> OrParams:
    d = dict(DEFAULTS)
    d.update(kw)
    if "horizon" not in d:
        d["horizon"] = d["dt"] * d["v_max"] * d["T"]          # MPPI_isaac.py:440 (float64 on host)
    if "target_speed" not in d:
        d["target_speed"] = d["v_max"]                         # MPPI_isaac.py:619
    p = OrParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


@dataclass
class StepResult:
    nominal1: np.ndarray
    nominal2: np.ndarray
    opt_v: np.ndarray
    opt_w: np.ndarray
    sim_traj: np.ndarray
    sim_heading: np.ndarray
    nominal1_f64: np.ndarray
    nominal2_f64: np.ndarray
    min_cost: float
    argmin: int
    weights_sum: float
    oob_clamps: int
    dump: dict = field(default_factory=dict)


_DUMP_SHAPES = {
    "u1": (1, np.float32), "u2": (1, np.float32), "v": (1, np.float32), "w": (1, np.float32),
    "traj": (3, np.float32), "heading": (3, np.float32), "lw": (3, np.float32), "rw": (3, np.float32),
    "dem_ij": (2, np.int32), "lw_ij": (2, np.int32), "rw_ij": (2, np.int32), "cm_ij": (2, np.int32),
}


def mppi_step(params: OrParams, dem: np.ndarray, half_width: float, costmap: np.ndarray, state: dict,
              nom1: np.ndarray, nom2: np.ndarray, eps1: np.ndarray, eps2: np.ndarray,
              dump: bool | list = False, nthreads: int = 1) -> StepResult:
    """One MPPI step on the CPU.  dem: (gs, gs) float32; costmap: (cms, cms) float32.
    state keys: x y hx hy hz wheel_l wheel_r sigma1 sigma2 goal_x goal_y goal_theta."""
    K, T = params.K, params.T
    dem = np.ascontiguousarray(dem, dtype=np.float32)
    costmap = np.ascontiguousarray(costmap, dtype=np.float32)
    gs, cms = dem.shape[0], costmap.shape[0]
    ter = OrTerrain(_p(dem), gs, half_width, 2.0 * half_width / gs, _p(costmap), cms, 2.0 * half_width / cms)
    st = OrState()
    for k, v in state.items():
        setattr(st, k, float(v))
    nom1 = np.ascontiguousarray(nom1, dtype=np.float32)
    nom2 = np.ascontiguousarray(nom2, dtype=np.float32)
    eps1 = np.ascontiguousarray(eps1, dtype=np.float32).reshape(K, T)
    eps2 = np.ascontiguousarray(eps2, dtype=np.float32).reshape(K, T)
    d = OrDump()
    keep = {}
    names = []
    if dump is True:
        names = list(_DUMP_SHAPES) + ["critics", "critics_ext", "cost", "weights"]
    elif dump:
        names = list(dump)
    for n in names:
        if n in _DUMP_SHAPES:
            c, dt = _DUMP_SHAPES[n]
            a = np.zeros((K, T) if c == 1 else (K, T, c), dtype=dt)
        elif n == "critics":
            a = np.zeros((K, 4), dtype=np.float32)
        elif n == "critics_ext":
            a = np.zeros((K, 6), dtype=np.float32)
        else:
            a = np.zeros(K, dtype=np.float32)
        keep[n] = a
        setattr(d, n, a.ctypes.data)
    o = OrOut()
    res = dict(nominal1=np.zeros(T, np.float32), nominal2=np.zeros(T, np.float32),
               opt_v=np.zeros(T, np.float32), opt_w=np.zeros(T, np.float32),
               sim_traj=np.zeros((T, 3), np.float32), sim_heading=np.zeros((T, 3), np.float32),
               nominal1_f64=np.zeros(T, np.float64), nominal2_f64=np.zeros(T, np.float64))
    for k, a in res.items():
        setattr(o, k, a.ctypes.data)
    rc = lib().oracle_mppi_step(C.byref(params), C.byref(ter), C.byref(st), _p(nom1), _p(nom2), _p(eps1), _p(eps2),
                                C.byref(d), C.byref(o), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_mppi_step failed: {rc}")
    return StepResult(min_cost=o.min_cost, argmin=o.argmin, weights_sum=o.weights_sum, oob_clamps=o.oob_clamps,
                      dump=keep, **res)


def combine_partials(parts: np.ndarray, T: int, lam: float, math: int = MATH_DET):
    parts = np.ascontiguousarray(parts, dtype=np.float32)
    G = parts.shape[0]
    n1 = np.zeros(T, np.float32)
    n2 = np.zeros(T, np.float32)
    m = np.zeros(1, np.float32)
    a = np.zeros(1, np.int32)
    s = np.zeros(1, np.float32)
    lib().oracle_combine_partials(_p(parts), G, T, lam, math, _p(n1), _p(n2), _p(m), _p(a), _p(s))
    return n1, n2, float(m[0]), int(a[0]), float(s[0])


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ------------------------------------------------------------------ obstacle costmap (MPPI_isaac.py:361-378)
def chamfer5x5(mask: np.ndarray) -> np.ndarray:
    """cv2.distanceTransform(mask, DIST_L2, 5) restated (two-pass 5x5 chamfer, 16.16 fixed point)."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.zeros(mask.shape, np.float32)
    rc = lib().oracle_chamfer5x5(mask.ctypes.data_as(C.c_void_p), mask.shape[0], mask.shape[1],
                                 out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle_chamfer5x5 failed")
    return out


def rasterize_obstacles(obstacles, origin, cms: int, half_width: float, r_robot: float, radius_scale: float = 0.5,
                        inflate: float = 0.1) -> np.ndarray:
    obs = np.ascontiguousarray(np.asarray(obstacles, np.float64).reshape(-1, 3))
    mask = np.zeros((cms, cms), np.uint8)
    f = lib().oracle_rasterize_obstacles
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                  C.c_double, C.c_void_p]
    rc = f(obs.ctypes.data_as(C.c_void_p), obs.shape[0], float(origin[0]), float(origin[1]), cms, float(half_width),
           float(r_robot), float(radius_scale), float(inflate), mask.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle_rasterize_obstacles failed")
    return mask


def obstacle_costmap(obstacles, origin, cms: int, half_width: float, r_robot: float, power: float = 20.0,
                     radius_scale: float = 0.5, inflate: float = 0.1):
    """Surface.create_obstacles_costmap restated: (mask, distance, costmap float32)."""
    mask = rasterize_obstacles(obstacles, origin, cms, half_width, r_robot, radius_scale, inflate)
    d = chamfer5x5(mask)
    smin, smax = float(d.min()), float(d.max())
    scale = (1.0 / (smax - smin)) if (smax - smin) > np.finfo(np.float64).eps else 0.0      # cv2.normalize, NORM_MINMAX
    shift = 0.0 - smin * scale
    dn = (d * np.float32(scale) + np.float32(shift)).astype(np.float32)
    return mask, d, ((np.float32(1) - dn) ** power).astype(np.float32)
