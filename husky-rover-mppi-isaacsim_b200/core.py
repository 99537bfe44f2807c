"""`Core`: object wrapper over one MppiHandle of libmppi_b200.so (include/mppi_b200.h).

This is the programmatic API below the reference-shaped `MPPI_Controller` facade; bench.py and the
multi-GPU drivers use it directly.  Device memory is held as torch tensors (plumbing only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import capi
from .devarray import view_device_memory


def make_state(x, y, heading=(1.0, 0.0, 0.0), wheel_l=0.0, wheel_r=0.0, sigma1=0.25, sigma2=0.25,
               goal_x=0.0, goal_y=0.0, goal_theta=2.2) -> capi.MppiState:
    h = np.asarray(heading, dtype=np.float64)
    h = h / np.linalg.norm(h)
    return capi.MppiState(float(x), float(y), float(h[0]), float(h[1]), float(h[2]), float(wheel_l), float(wheel_r),
                          float(sigma1), float(sigma2), float(goal_x), float(goal_y), float(goal_theta))


class Core:
    def __init__(self, K: int, T: int, device: int = 0, math: str = "strict", max_rovers: int = 1,
                 params: Optional[capi.MppiParams] = None, **overrides):
        if not torch.cuda.is_available():
            raise capi.MppiError("the MPPI core needs a CUDA device (there is no CPU fallback)")
        self.L = capi.lib()
        p = params if params is not None else capi.default_params(K, T)
        p.K, p.T = K, T
        p.math = capi.MATH_STRICT if math == "strict" else capi.MATH_FAST
        for k, v in overrides.items():
            setattr(p, "lam" if k in ("lambda_", "temperature") else k, v)
        # horizon = dt * v_max * T is formed in double on the host (MPPI_isaac.py:440), then cast to fp32
        if params is None and "horizon" not in overrides:
            p.horizon = overrides.get("dt", 0.045) * overrides.get("v_max", 2.0) * T
        if params is None and "target_speed" not in overrides:
            p.target_speed = overrides.get("v_max", 2.0)
        self.p, self.K, self.T = p, K, T
        self.device = torch.device("cuda", device)
        self.max_rovers = max_rovers
        self.h = C.c_void_p()
        capi.check(self.L.mppi_create(C.byref(p), device, max_rovers, C.byref(self.h)), "mppi_create")
        out = capi.MppiOutputs()
        capi.check(self.L.mppi_get_outputs(self.h, C.byref(out)), "mppi_get_outputs")
        R, dev = max_rovers, self.device
        self.optimal_u1 = view_device_memory(out.optimal_u1, (R, T), dev)
        self.optimal_u2 = view_device_memory(out.optimal_u2, (R, T), dev)
        self.optimal_v = view_device_memory(out.optimal_v, (R, T), dev)
        self.optimal_w = view_device_memory(out.optimal_w, (R, T), dev)
        self.costs = view_device_memory(out.costs, (R, K), dev)
        self.stats = view_device_memory(out.stats, (R, capi.STATS_STRIDE), dev)
        self.sim_traj = view_device_memory(out.sim_traj, (T, 3), dev)
        self.sim_heading = view_device_memory(out.sim_heading, (T, 3), dev)
        self._keep = {}
        self._cmd = (C.c_float * 2)()

    # ------------------------------------------------------------------ terrain
    def set_terrain(self, dem: torch.Tensor, half_width: float, costmap: torch.Tensor):
        """dem [gs, gs], costmap [cms, cms]: float32 CUDA tensors (borrowed, kept alive by this object)."""
        assert dem.is_cuda and costmap.is_cuda and dem.dtype == torch.float32 and costmap.dtype == torch.float32
        dem, costmap = dem.contiguous(), costmap.contiguous()
        gs, cms = dem.shape[0], costmap.shape[0]
        self._keep["terrain"] = (dem, costmap)
        t = capi.MppiTerrain(dem.data_ptr(), gs, half_width, 2.0 * half_width / gs,
                             costmap.data_ptr(), cms, 2.0 * half_width / cms)
        capi.check(self.L.mppi_set_terrain(self.h, C.byref(t)), "mppi_set_terrain")

    def set_terrain_batched(self, dems: torch.Tensor, half_width: float, costmaps: torch.Tensor):
        """dems [R, gs, gs], costmaps [R, cms, cms]: one map per rover."""
        R, gs, cms = dems.shape[0], dems.shape[1], costmaps.shape[1]
        dems, costmaps = dems.contiguous(), costmaps.contiguous()
        arr = (capi.MppiTerrain * R)()
        for r in range(R):
            arr[r] = capi.MppiTerrain(dems[r].data_ptr(), gs, half_width, 2.0 * half_width / gs,
                                      costmaps[r].data_ptr(), cms, 2.0 * half_width / cms)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        self._keep["terrains"] = (dems, costmaps, raw)
        capi.check(self.L.mppi_set_terrain_batched(self.h, raw.data_ptr(), R), "mppi_set_terrain_batched")

    def set_terrain_batched_shared(self, dems: torch.Tensor, half_width: float, costmaps: torch.Tensor):
        """Batched controllers over ONE shared map: dems / costmaps are [R, gs, gs] / [R, cms, cms] VIEWS (expanded,
        stride 0 over rovers) of a single DEM / costmap -- every rover's MppiTerrain points at the same memory."""
        R, gs, cms = dems.shape[0], dems.shape[1], costmaps.shape[1]
        arr = (capi.MppiTerrain * R)()
        for r in range(R):
            arr[r] = capi.MppiTerrain(dems[r].data_ptr(), gs, half_width, 2.0 * half_width / gs,
                                      costmaps[r].data_ptr(), cms, 2.0 * half_width / cms)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        self._keep["terrains"] = (dems, costmaps, raw)
        capi.check(self.L.mppi_set_terrain_batched(self.h, raw.data_ptr(), R), "mppi_set_terrain_batched")

    @staticmethod
    def pack_states(states, device) -> torch.Tensor:
        arr = (capi.MppiState * len(states))(*states)
        return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)

    # ------------------------------------------------------------------ stepping
    def _stream(self, stream):
        return (stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream

    def step(self, state: capi.MppiState, proj: int = capi.PROJ_3D, noise: Optional[torch.Tensor] = None,
             seed: int = 0, offset: int = 0, stream=None):
        capi.check(self.L.mppi_step(self.h, C.byref(state), proj, noise.data_ptr() if noise is not None else None,
                                    seed, offset, self._stream(stream)), "mppi_step")

    def step_host(self, state: capi.MppiState, proj: int = capi.PROJ_3D, seed: int = 0, offset: int = 0, stream=None):
        """Host state in, host command (v, w) out: includes the D2H of the result and the synchronisation."""
        capi.check(self.L.mppi_step_host(self.h, C.byref(state), proj, seed, offset, self._cmd, self._stream(stream)),
                   "mppi_step_host")
        return self._cmd[0], self._cmd[1]

    def step_batched(self, states_dev: torch.Tensor, n_rovers: int, proj: int = capi.PROJ_3D, seed: int = 0,
                     offset: int = 0, stream=None):
        capi.check(self.L.mppi_step_batched(self.h, states_dev.data_ptr(), n_rovers, proj, seed, offset,
                                            self._stream(stream)), "mppi_step_batched")

    def step_partial(self, state, partial_out: torch.Tensor, k_begin: int, proj: int = capi.PROJ_3D,
                     noise: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0, stream=None):
        capi.check(self.L.mppi_step_partial(self.h, C.byref(state), proj,
                                            noise.data_ptr() if noise is not None else None, seed, offset, k_begin,
                                            partial_out.data_ptr(), self._stream(stream)), "mppi_step_partial")

    def combine_partials(self, state, parts: torch.Tensor, n_parts: int, stream=None):
        capi.check(self.L.mppi_combine_partials(self.h, C.byref(state), parts.data_ptr(), n_parts,
                                                self._stream(stream)), "mppi_combine_partials")

    # ---- sample-sharded step with the exchange fused into the launch (NVLink peer memory, CUDA IPC)
    def comm_export(self, world: int) -> bytes:
        buf = (C.c_ubyte * 64)()
        capi.check(self.L.mppi_comm_export(self.h, world, buf), "mppi_comm_export")
        return bytes(buf)

    def comm_connect(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        capi.check(self.L.mppi_comm_connect(self.h, rank, world, buf), "mppi_comm_connect")

    def step_sharded(self, state, k_begin: int, proj: int = capi.PROJ_3D, noise: Optional[torch.Tensor] = None,
                     seed: int = 0, offset: int = 0, stream=None):
        capi.check(self.L.mppi_step_sharded(self.h, C.byref(state), proj,
                                            noise.data_ptr() if noise is not None else None, seed, offset, k_begin,
                                            self._stream(stream)), "mppi_step_sharded")

    def step_sharded_host(self, state, k_begin: int, proj: int = capi.PROJ_3D, seed: int = 0, offset: int = 0,
                          stream=None):
        capi.check(self.L.mppi_step_sharded_host(self.h, C.byref(state), proj, seed, offset, k_begin, self._cmd,
                                                 self._stream(stream)), "mppi_step_sharded_host")
        return self._cmd[0], self._cmd[1]

    def run_closed_loop(self, state: capi.MppiState, max_iters: int, proj: int = capi.PROJ_3D, seed: int = 0,
                        offset0: int = 0, goal_tol: float = 0.5, sigma_base: float = 0.4, sigma_gain: float = 1.0,
                        noise: Optional[torch.Tensor] = None, want_log: bool = True, stream=None):
        """Device-resident closed loop (mppi_run_closed_loop).  Returns (iterations done, goal reached, log[k, 8]);
        `state` is updated in place to the state after the last iteration."""
        log = np.zeros((max_iters, 8), np.float32) if want_log else None
        done, reached = C.c_int32(0), C.c_int32(0)
        capi.check(self.L.mppi_run_closed_loop(self.h, C.byref(state), proj,
                                               noise.data_ptr() if noise is not None else None, seed, offset0,
                                               max_iters, goal_tol, sigma_base, sigma_gain,
                                               log.ctypes.data if want_log else None, C.byref(done),
                                               C.byref(reached), self._stream(stream)), "mppi_run_closed_loop")
        return int(done.value), bool(reached.value), (log[:done.value] if want_log else None)

    def sim_rollout(self, state, stream=None):
        capi.check(self.L.mppi_sim_rollout(self.h, C.byref(state), self._stream(stream)), "mppi_sim_rollout")

    def enable_timing(self, on: bool = True):
        capi.check(self.L.mppi_enable_timing(self.h, 1 if on else 0), "mppi_enable_timing")

    def latency_stats(self) -> dict:
        """p50 / p99 / max device time (us) of the last up-to-1024 completed steps since enable_timing()."""
        a, b, c, n = C.c_float(), C.c_float(), C.c_float(), C.c_int32()
        capi.check(self.L.mppi_latency_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)), "mppi_latency_stats")
        return dict(p50_us=float(a.value), p99_us=float(b.value), max_us=float(c.value), n=int(n.value))

    def partial_floats(self) -> int:
        return int(self.L.mppi_partial_floats(self.T))

    def set_nominal(self, u1: np.ndarray, u2: np.ndarray, n_rovers: int = 1):
        u1 = np.ascontiguousarray(u1, np.float32)
        u2 = np.ascontiguousarray(u2, np.float32)
        capi.check(self.L.mppi_set_nominal(self.h, u1.ctypes.data, u2.ctypes.data, n_rovers, self._stream(None)),
                   "mppi_set_nominal")

    def read_stats(self, rover: int = 0) -> dict:
        s = self.stats[rover].cpu().numpy()
        i = s.view(np.int32)
        return dict(min_cost=float(s[0]), argmin=int(i[1]), weights_sum=float(s[2]), oob=int(i[3]), nan=int(i[4]),
                    ess=float(s[5]), v0=float(s[6]), w0=float(s[7]))

    def close(self):
        if getattr(self, "h", None):
            self.L.mppi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
