"""Shared helpers of the test-suite: scenarios, oracle runs, and GPU runs through the C ABI."""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np


@functools.lru_cache(maxsize=8)
def terrain(name: str = "C1"):
    """(dem [gs,gs] f32, costmap [cms,cms] f32, half_width) of a BASELINE workload, or small test maps."""
    from mppi_b200 import synthetic as syn
    if name == "small":                       # 25.6 m map, 0.1 m DEM, 0.2 m costmap, bumpy, dense rocks
        hw, gs, cms = 12.8, 256, 128
        bumps = [((-3.0, 2.0), 1.4, 2.5), ((4.0, -1.0), 2.0, 3.0), ((0.5, 6.0), 0.9, 1.5), ((-6.0, -6.0), 1.2, 2.0)]
        dem = syn.crater_dem(gs, hw, bumps=bumps).numpy()
        cm = syn.rock_costmap(cms, hw, n_rocks=60, seed=5)
        return dem, cm, hw
    if name == "rough":                       # the C1 map with 2 cm of cell-to-cell noise: slopes change by several
        dem, cm, hw = terrain("C1")           # degrees per step, as on a photogrammetric DEM
        rng = np.random.default_rng(2024)
        return (dem + rng.normal(0.0, 0.02, dem.shape).astype(np.float32)), cm, hw
    w = syn.WORKLOADS[name]
    dem = syn.crater_dem(w.grid_size, w.half_width).numpy()
    cm = syn.rock_costmap(w.costmap_size, w.half_width)
    return dem, cm, w.half_width


def default_state(**kw):
    st = dict(x=-60.57, y=-60.23, hx=1.0, hy=0.0, hz=0.0, wheel_l=0.0, wheel_r=0.0, sigma1=0.25, sigma2=0.25,
              goal_x=65.8, goal_y=65.4, goal_theta=2.2)
    st.update(kw)
    return st


def normals(K, T, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((K, T)).astype(np.float32), rng.standard_normal((K, T)).astype(np.float32))


# ------------------------------------------------------------------ GPU side, through the C ABI
class GpuCore:
    """Thin test driver over libmppi_b200.so (ctypes + torch for device memory)."""

    def __init__(self, K, T, dem, cm, hw, math="strict", max_rovers=1, **param_overrides):
        import torch
        from mppi_b200 import capi
        self.torch, self.capi = torch, capi
        self.L = capi.lib()
        self.K, self.T = K, T
        p = capi.default_params(K, T)
        p.math = capi.MATH_STRICT if math == "strict" else capi.MATH_FAST
        for k, v in param_overrides.items():
            setattr(p, "lam" if k == "lambda_" else k, v)
        # horizon = dt * v_max * T is formed in DOUBLE on the host (MPPI_isaac.py:440) and only then cast to fp32
        if "horizon" not in param_overrides:
            p.horizon = param_overrides.get("dt", 0.045) * param_overrides.get("v_max", 2.0) * T
        if "target_speed" not in param_overrides:
            p.target_speed = param_overrides.get("v_max", 2.0)
        self.p = p
        self.h = C.c_void_p()
        capi.check(self.L.mppi_create(C.byref(p), 0, max_rovers, C.byref(self.h)), "create")
        self.dev = torch.device("cuda", 0)
        self.dem = torch.from_numpy(np.ascontiguousarray(dem, np.float32)).to(self.dev)
        self.cm = torch.from_numpy(np.ascontiguousarray(cm, np.float32)).to(self.dev)
        gs, cms = dem.shape[0], cm.shape[0]
        self.ter = capi.MppiTerrain(self.dem.data_ptr(), gs, hw, 2.0 * hw / gs, self.cm.data_ptr(), cms, 2.0 * hw / cms)
        capi.check(self.L.mppi_set_terrain(self.h, C.byref(self.ter)), "set_terrain")
        self.out = capi.MppiOutputs()
        capi.check(self.L.mppi_get_outputs(self.h, C.byref(self.out)), "outputs")

    def state(self, st: dict):
        s = self.capi.MppiState()
        for k, v in st.items():
            setattr(s, k, float(v))
        return s

    def set_nominal(self, n1, n2):
        n1 = np.ascontiguousarray(n1, np.float32)
        n2 = np.ascontiguousarray(n2, np.float32)
        self.capi.check(self.L.mppi_set_nominal(self.h, n1.ctypes.data, n2.ctypes.data, 1, None), "set_nominal")

    def get_nominal(self):
        n1 = np.zeros(self.T, np.float32)
        n2 = np.zeros(self.T, np.float32)
        self.capi.check(self.L.mppi_get_nominal(self.h, n1.ctypes.data, n2.ctypes.data, 1, None), "get_nominal")
        return n1, n2

    def view(self, ptr, shape, dtype=None):
        from mppi_b200.devarray import view_device_memory
        return view_device_memory(ptr, shape, self.dev, dtype or self.torch.float32)

    def step(self, st: dict, proj=3, eps=None, seed=0, offset=0):
        """eps: None (Philox) or (eps1, eps2) numpy [K,T]."""
        torch = self.torch
        noise = None
        if eps is not None:
            self._noise = torch.from_numpy(np.stack([eps[0], eps[1]]).astype(np.float32)).to(self.dev).contiguous()
            noise = self._noise.data_ptr()
        s = self.state(st)
        self.capi.check(self.L.mppi_step(self.h, C.byref(s), proj, noise, seed, offset, None), "step")
        torch.cuda.synchronize()
        T, K = self.T, self.K
        res = dict(
            nominal1=self.view(self.out.optimal_u1, (T,)).cpu().numpy(),
            nominal2=self.view(self.out.optimal_u2, (T,)).cpu().numpy(),
            opt_v=self.view(self.out.optimal_v, (T,)).cpu().numpy(),
            opt_w=self.view(self.out.optimal_w, (T,)).cpu().numpy(),
            cost=self.view(self.out.costs, (K,)).cpu().numpy(),
        )
        stats = self.view(self.out.stats, (8,)).cpu().numpy()
        si = stats.view(np.int32)
        res.update(min_cost=float(stats[0]), argmin=int(si[1]), weights_sum=float(stats[2]), oob=int(si[3]),
                   nan=int(si[4]), ess=float(stats[5]), v0=float(stats[6]), w0=float(stats[7]))
        return res

    def sim_rollout(self, st: dict):
        s = self.state(st)
        self.capi.check(self.L.mppi_sim_rollout(self.h, C.byref(s), None), "sim")
        self.torch.cuda.synchronize()
        return (self.view(self.out.sim_traj, (self.T, 3)).cpu().numpy(),
                self.view(self.out.sim_heading, (self.T, 3)).cpu().numpy())

    def dump(self, st: dict, proj=3, eps=None, seed=0, offset=0, previous=True, names=None):
        torch, capi = self.torch, self.capi
        K, T = self.K, self.T
        shapes = {"u1": (K, T), "u2": (K, T), "v": (K, T), "w": (K, T), "traj": (K, T, 3), "heading": (K, T, 3),
                  "lw": (K, T, 3), "rw": (K, T, 3), "dem_ij": (K, T, 2), "lw_ij": (K, T, 2), "rw_ij": (K, T, 2),
                  "cm_ij": (K, T, 2), "critics": (K, 4), "weights": (K,), "critics_ext": (K, 6)}
        names = names or list(shapes)
        d = capi.MppiDebugDump()
        out = {}
        for n in names:
            out[n] = torch.zeros(shapes[n], dtype=torch.int32 if n.endswith("_ij") else torch.float32, device=self.dev)
            setattr(d, n, out[n].data_ptr())
        noise = None
        if eps is not None:
            self._noise = torch.from_numpy(np.stack([eps[0], eps[1]]).astype(np.float32)).to(self.dev).contiguous()
            noise = self._noise.data_ptr()
        s = self.state(st)
        capi.check(self.L.mppi_debug_dump(self.h, C.byref(s), proj, noise, seed, offset, 1 if previous else 0,
                                          C.byref(d), None), "dump")
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    def close(self):
        if self.h:
            self.L.mppi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
