"""ctypes binding of libmppi_b200.so (include/mppi_b200.h).

The library is the product: if it is missing or fails to load this module raises -- there is no
CPU or PyTorch fallback for the MPPI step.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPPI_B200_LIB") or os.path.join(_HERE, "libmppi_b200.so")   # override: A/B builds
CSRC = os.path.join(_HERE, "csrc")

MPPI_OK = 0
ABI_VERSION = 4
PROJ_2D = 2
PROJ_3D = 3
MATH_STRICT = 0
MATH_FAST = 1
VARIANT_AUTO, VARIANT_MONO, VARIANT_PIPE = 0, 1, 2
INPUT_SKID_STEER, INPUT_UNICYCLE = 0, 1
PARTIAL_HEADER = 4
STATS_STRIDE = 8


class MppiParams(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("T", C.c_int32), ("math", C.c_int32), ("variant", C.c_int32),
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
        # optional critics (weight 0 = off): dormant reference critics, then the roll / pitch / effort extensions
        ("cw_orient", C.c_float), ("cw_slope_path", C.c_float), ("cw_goal_angle", C.c_float),
        ("goal_angle_radius", C.c_float), ("cw_roll", C.c_float), ("cw_pitch", C.c_float), ("cw_effort", C.c_float),
        ("reserved", C.c_int32),
    ]


class MppiTerrain(C.Structure):
    _fields_ = [
        ("dem", C.c_void_p), ("grid_size", C.c_int32), ("half_width", C.c_float), ("resolution", C.c_float),
        ("costmap", C.c_void_p), ("costmap_size", C.c_int32), ("costmap_resolution", C.c_float),
    ]


class MppiState(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("x", "y", "hx", "hy", "hz", "wheel_l", "wheel_r", "sigma1", "sigma2",
                 "goal_x", "goal_y", "goal_theta")]


class MppiOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("optimal_u1", "optimal_u2", "optimal_v", "optimal_w", "costs", "stats", "sim_traj", "sim_heading")]


class MppiDebugDump(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("u1", "u2", "v", "w", "traj", "heading", "lw", "rw",
                 "dem_ij", "lw_ij", "rw_ij", "cm_ij", "critics", "weights", "critics_ext")]


# every symbol include/mppi_b200.h declares: (name, restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "mppi_strerror": (C.c_char_p, [C.c_int]),
    "mppi_abi_version": (C.c_int, []),
    "mppi_default_params": (C.c_int, [C.POINTER(MppiParams), C.c_int32, C.c_int32]),
    "mppi_create": (C.c_int, [C.POINTER(MppiParams), C.c_int32, C.c_int32, C.POINTER(_H)]),
    "mppi_destroy": (C.c_int, [_H]),
    "mppi_set_params": (C.c_int, [_H, C.POINTER(MppiParams)]),
    "mppi_set_terrain": (C.c_int, [_H, C.POINTER(MppiTerrain)]),
    "mppi_set_terrain_batched": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "mppi_set_nominal": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_get_nominal": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_step": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mppi_step_host": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mppi_step_batched": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mppi_step_partial": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64,
                                    C.c_uint32, C.c_void_p, C.c_void_p]),
    "mppi_combine_partials": (C.c_int, [_H, C.POINTER(MppiState), C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_partial_floats": (C.c_int, [C.c_int32]),
    "mppi_comm_export": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "mppi_comm_connect": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
    "mppi_step_sharded_host": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32,
                                         C.c_void_p, C.c_void_p]),
    "mppi_step_sharded": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64,
                                    C.c_uint32, C.c_void_p]),
    "mppi_run_closed_loop": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64,
                                       C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]),
    "mppi_closed_loop_last_input": (C.c_int, [_H, C.POINTER(MppiState)]),
    "mppi_sim_rollout": (C.c_int, [_H, C.POINTER(MppiState), C.c_void_p]),
    "mppi_debug_dump": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32,
                                  C.POINTER(MppiDebugDump), C.c_void_p]),
    "mppi_build_costmap": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "mppi_export_trajectories": (C.c_int, [_H, C.POINTER(MppiState), C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mppi_get_outputs": (C.c_int, [_H, C.POINTER(MppiOutputs)]),
    "mppi_enable_timing": (C.c_int, [_H, C.c_int32]),
    "mppi_last_step_us": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "mppi_latency_stats": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_int32)]),
    "mppi_set_trace": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_int32)]),
    "mppi_measure_peaks": (C.c_int, [C.c_int32, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float)]),
    "mppi_test_timestamp": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_test_detmath": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_test_noise": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppi_test_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mppi_test_divsqrt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}


class MppiError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libmppi_b200.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "mppi_b200.h"))
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    r = subprocess.run(["make", "-j4", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise MppiError("building libmppi_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raises MppiError if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MppiError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(the MPPI core has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        ab = bool(os.environ.get("MPPI_B200_LIB"))       # A/B against an older build: newer entry points may be absent
        for name, (res, args) in SYMBOLS.items():
            if ab and not hasattr(L, name):
                continue
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.mppi_abi_version() != ABI_VERSION and not ab:
            raise MppiError("libmppi_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != MPPI_OK:
        msg = lib().mppi_strerror(rc).decode()
        raise MppiError(f"{what}: {msg} (status {rc})")


def default_params(K: int, T: int) -> MppiParams:
    p = MppiParams()
    check(lib().mppi_default_params(C.byref(p), K, T), "mppi_default_params")
    return p
