"""Host-side mirror of the reference controller interface (drop-in boundary).

Mirrors thesis_master/warp_implementation/MPPI_isaac.py: `Surface` (:259-378), `Robot` (:381-400) and
`MPPI_Controller` (:402-805) keep their constructor signatures, method names (`warp_setup`, `reset`,
`MPPI_step`, `run`) and the attributes the Isaac driver reads and writes between steps
(visual_terrain_stack_full_terrain.py:466-576).  Where the reference issues nine Warp launches per
iteration, `MPPI_step` makes ONE call into libmppi_b200.so.  There is no CPU fallback: without the
CUDA library and a GPU the step raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import yaml

from . import capi
from .devarray import DeviceArray, to_device_f32, view_device_memory


# --------------------------------------------------------------------------------------------- Surface
class Surface:
    """DEM + obstacle costmap container, MPPI_isaac.py:259-378 (same constructor signature)."""

    def __init__(self, which_map, filename, which_costmap, costmap_file, grid_size, half_width, origin, bumps,
                 radius_robot, obstacles=[]):
        self.grid_size = grid_size
        self.r_robot = radius_robot
        self.half_width = half_width
        self.resolution = 2 * self.half_width / self.grid_size                 # MPPI_isaac.py:265
        self.costmap_size = int(self.grid_size / 8)                            # MPPI_isaac.py:271
        self.costmap_resolution = 2 * self.half_width / self.costmap_size      # MPPI_isaac.py:272
        self.Z = np.zeros((grid_size, grid_size), dtype=np.float32)
        self.costmap = np.zeros((self.costmap_size, self.costmap_size), dtype=np.float32)
        self.obstacles = obstacles
        self._grids = {}
        if which_map == "manual":
            self.Z = self.create_surface(bumps)
        if which_map == "imported":
            _, _, self.Z = self.import_surface(filename, 1000, 2500, bumps)      # MPPI_isaac.py:285-288
        if which_costmap == "manual":
            self.costmap = self.create_obstacles_costmap(obstacles, origin)
        if which_map == "imported":                                              # MPPI_isaac.py:296-297 (unconditional)
            self.costmap = self.import_obstacles_costmap(costmap_file)

    # The reference builds four full meshgrids in its constructor (MPPI_isaac.py:267-277: X, Y over the DEM grid,
    # X_costmap, Y_costmap over the costmap grid; 7000^2 float64 x 2 = 784 MB at the Isaac sizes).  Only its plotting
    # helpers read them, so here they are built on first access.
    def _grid(self, name):
        if name not in self._grids:
            n = self.grid_size if name in ("X", "Y") else self.costmap_size
            ax = np.linspace(-self.half_width, self.half_width, n)
            gx, gy = np.meshgrid(ax, ax)
            self._grids["X" if name in ("X", "Y") else "X_costmap"] = gx
            self._grids["Y" if name in ("X", "Y") else "Y_costmap"] = gy
        return self._grids[name]

    X = property(lambda self: self._grid("X"))
    Y = property(lambda self: self._grid("Y"))
    X_costmap = property(lambda self: self._grid("X_costmap"))
    Y_costmap = property(lambda self: self._grid("Y_costmap"))

    def import_surface(self, filename, start_index, end_index, bumps):
        """MPPI_isaac.py:301-307: returns (X, Y, Z) with Z loaded from `filename` (X, Y span end - start samples)."""
        ax = np.linspace(-self.half_width, self.half_width, end_index - start_index)
        X, Y = np.meshgrid(ax, ax)
        return X, Y, np.load(filename)

    def import_obstacles_costmap(self, costmap_file):
        """MPPI_isaac.py:358-359."""
        return np.load(costmap_file)

    def create_surface(self, bumps) -> np.ndarray:
        """Crater field: sum of (h-0.5) exp(-r^2/2w^2) - (h+0.5) exp(-r^2/2(w/2)^2), MPPI_isaac.py:317-320."""
        x = np.linspace(-self.half_width, self.half_width, self.grid_size)
        X, Y = np.meshgrid(x, x)
        Z = np.zeros_like(X)
        for (cx, cy), h, w in bumps:
            r2 = (X - cx) ** 2 + (Y - cy) ** 2
            Z += (h - 0.5) * np.exp(-r2 / (2 * w ** 2))
            Z -= (h + 0.5) * np.exp(-r2 / (2 * (w / 2) ** 2))
        return Z

    def create_obstacles_costmap(self, obstacles, origin, power: float = 20.0) -> np.ndarray:
        """Disc rasterisation + L2 distance transform + (1-d)^20, MPPI_isaac.py:361-378."""
        import cv2
        n = self.costmap_size
        xc = np.linspace(-self.half_width, self.half_width, n)
        Xc, Yc = np.meshgrid(xc, xc)
        free = 255 * np.ones((n, n), dtype=np.uint8)
        x0, y0 = origin
        for x_global, y_global, r_obs in obstacles:
            x_local = y_global - y0
            y_local = x_global - x0
            total_radius = r_obs / 2 + self.r_robot + 0.1
            free[(Xc - x_local) ** 2 + (Yc - y_local) ** 2 <= total_radius ** 2] = 0
        dist = cv2.distanceTransform(free, cv2.DIST_L2, 5)
        dist = cv2.normalize(dist, None, 0, 1.0, cv2.NORM_MINMAX)
        return (1 - dist) ** power


# --------------------------------------------------------------------------------------------- Robot
class Robot:
    """Pose history + wheel speeds, MPPI_isaac.py:381-400."""

    def __init__(self, x, y, heading_vector, config_file):
        with open(config_file, "r") as f:
            config = yaml.safe_load(f)
        self.x = [x]
        self.y = [y]
        self.z = [0]
        self.lin_vel = []
        self.ang_vel = []
        self.heading_vector = np.array(heading_vector, dtype=np.float64) / np.linalg.norm(heading_vector)
        self.radius = config["frame_work"]["robot_radius"]
        self.left_wheel_speed = 0.0
        self.right_wheel_speed = 0.0

    def update_position(self, new_x, new_y, new_z, new_heading):
        self.x.append(new_x)
        self.y.append(new_y)
        self.z.append(new_z)
        self.heading_vector = new_heading


# --------------------------------------------------------------------------------------------- Controller
class MPPI_Controller:
    """Drop-in for MPPI_isaac.MPPI_Controller backed by the fused sm_100a kernel.

    Extra keyword arguments (all optional, defaults reproduce the reference behaviour):
      math    "strict" (bit-reproducible vs the oracle) | "fast"
      device  CUDA device index
      seed    Philox seed (the reference seeds its host generator with 42, MPPI_isaac.py:409)
      critic_weights  dict of critic weights (MppiParams.cw_*), e.g. {"cw_orient": 1.0, "cw_slope_path": 50.5} to
              re-enable the critics the reference keeps commented out (critics_warp.py:324,326), "cw_goal_angle"
              (critics_warp.py:5-41) or the roll / pitch / effort extensions.  Empty = the reference's four critics
              with its weights.
    """

    def __init__(self, surface, robot, config_path, goal_x, goal_y, goal_orientation, *, math: str = "strict",
                 device: int = 0, seed: int = 42, overrides: Optional[dict] = None,
                 critic_weights: Optional[dict] = None):
        with open(config_path, "r") as f:
            config = yaml.safe_load(f)
        self.rng = np.random.default_rng(seed=42)
        self.robot = robot
        self.surface = surface
        self.goal_x = goal_x
        self.goal_y = goal_y
        self.goal = (goal_x, goal_y)
        self.goal_orientation = goal_orientation
        self.loop = 0

        c = config["controller"]
        self.number_of_iterations = int(c["number_of_iterations"])
        self.dt = float(c["dt"])
        self.number_of_trajectories = int(c["number_of_trajectories"])
        v, i = config["velocities"], config["inputs"]
        self.initial_linear_velocity = float(v["initial_linear_velocity"])
        self.initial_angular_velocity = float(v["initial_angular_velocity"])
        self.v_min_linear, self.v_max_linear = float(v["min_linear_velocity"]), float(v["max_linear_velocity"])
        self.v_min_angular, self.v_max_angular = float(v["min_angular_velocity"]), float(v["max_angular_velocity"])
        self.std_dev_u1, self.std_dev_u2 = float(i["std_dev_u1"]), float(i["std_dev_u2"])
        self.min_u1, self.max_u1 = float(i["min_u1"]), float(i["max_u1"])
        self.min_u2, self.max_u2 = float(i["min_u2"]), float(i["max_u2"])
        self.temperature = float(config["cost_evaluation"]["temperature"])
        if overrides:
            for k, val in overrides.items():
                setattr(self, k, val)
        self.horizon = self.dt * self.v_max_linear * self.number_of_iterations     # MPPI_isaac.py:440

        self.math = {"strict": capi.MATH_STRICT, "fast": capi.MATH_FAST}[math]
        self.critic_weights = dict(critic_weights or {})
        for k in self.critic_weights:
            if not (k.startswith("cw_") or k == "goal_angle_radius") or not hasattr(capi.MppiParams, k):
                raise ValueError(f"unknown critic weight {k!r}")
        self.device = torch.device("cuda", device)
        self.seed = int(seed)
        self._handle = None
        self._step_count = 0
        self._sim_stale = True
        self._last_state = None
        self._last_proj = capi.PROJ_3D
        self._last_offset = 0
        self._injected_noise = None

    # ------------------------------------------------------------------ setup
    def _params(self) -> capi.MppiParams:
        p = capi.default_params(self.number_of_trajectories, self.number_of_iterations)
        p.math = self.math
        p.dt = self.dt
        p.u1_min, p.u1_max, p.u2_min, p.u2_max = self.min_u1, self.max_u1, self.min_u2, self.max_u2
        p.v_min, p.v_max = self.v_min_linear, self.v_max_linear
        p.w_min, p.w_max = self.v_min_angular, self.v_max_angular
        p.lam = self.temperature
        p.r_wheels = float(self.robot.radius)
        p.horizon = self.horizon
        p.target_speed = self.v_max_linear
        for k, val in self.critic_weights.items():
            setattr(p, k, float(val))
        return p

    def warp_setup(self):
        """Allocates device state and uploads the terrain (replaces MPPI_isaac.py:442-487).  The name is kept
        for drop-in compatibility; nothing here uses Warp."""
        if not torch.cuda.is_available():
            raise capi.MppiError("MPPI_Controller needs a CUDA device (no CPU fallback)")
        L = capi.lib()
        self._lib = L
        if self._handle is not None:
            L.mppi_destroy(self._handle)
        h = C.c_void_p()
        self._p = self._params()
        capi.check(L.mppi_create(C.byref(self._p), self.device.index, 1, C.byref(h)), "mppi_create")
        self._handle = h
        T, K = self.number_of_iterations, self.number_of_trajectories
        out = capi.MppiOutputs()
        capi.check(L.mppi_get_outputs(h, C.byref(out)), "mppi_get_outputs")
        dev = self.device
        self.optimal_u1_wp = DeviceArray(view_device_memory(out.optimal_u1, (T,), dev))
        self.optimal_u2_wp = DeviceArray(view_device_memory(out.optimal_u2, (T,), dev))
        self.optimal_lin_vel_wp = DeviceArray(view_device_memory(out.optimal_v, (T,), dev))
        self.optimal_ang_vel_wp = DeviceArray(view_device_memory(out.optimal_w, (T,), dev))
        self.optimal_lin_vel_wp.tensor.fill_(self.initial_linear_velocity)       # MPPI_isaac.py:453-454
        self.optimal_ang_vel_wp.tensor.fill_(self.initial_angular_velocity)
        self.costs_wp = DeviceArray(view_device_memory(out.costs, (K,), dev))
        self._stats = view_device_memory(out.stats, (capi.STATS_STRIDE,), dev)
        self.trajectories_sim = DeviceArray(view_device_memory(out.sim_traj, (T, 3), dev), on_read=self._ensure_sim)
        self.heading_vectors_sim = DeviceArray(view_device_memory(out.sim_heading, (T, 3), dev),
                                               on_read=self._ensure_sim)
        self._Z = to_device_f32(self.surface.Z, dev)
        self.costmap_wp = DeviceArray(to_device_f32(self.surface.costmap, dev))
        self._terrain_dirty = True
        self._cmd = (C.c_float * 2)()
        self._stream = torch.cuda.current_stream(dev).cuda_stream
        torch.cuda.synchronize(dev)

    setup = warp_setup

    # Z_wp can be re-pointed at another device array between steps (driver :567): zero-copy.
    @property
    def Z_wp(self):
        return DeviceArray(self._Z)

    @Z_wp.setter
    def Z_wp(self, arr):
        self._Z = to_device_f32(arr, self.device)
        self._terrain_dirty = True

    def rebuild_costmap(self, obstacles, origin, power: float = 20.0):
        """Block-change path of the Isaac driver (visual_terrain_stack_full_terrain.py:561-563:
        `surface.costmap = surface.create_obstacles_costmap(...)`; `costmap_wp.assign(surface.costmap.flatten())`) done
        on the device: rocks -> distance transform -> (1 - d)^p written straight into `costmap_wp`.  `surface.costmap`
        is refreshed lazily from the device copy only if somebody reads it (`surface_costmap()`)."""
        from .costmap import build_obstacle_costmap
        s = self.surface
        s.obstacles = obstacles
        build_obstacle_costmap(obstacles, origin, int(s.costmap_size), float(s.half_width), float(s.r_robot),
                               out=self.costmap_wp.tensor.view(int(s.costmap_size), int(s.costmap_size)),
                               device=self.device.index, power=power)
        self._terrain_dirty = True

    def surface_costmap(self) -> np.ndarray:
        """Host copy of the device costmap (refreshes `surface.costmap`)."""
        n = int(self.surface.costmap_size)
        self.surface.costmap = self.costmap_wp.tensor.view(n, n).cpu().numpy()
        return self.surface.costmap

    def _push_terrain(self):
        s = self.surface
        gs, cms = int(s.grid_size), int(s.costmap_size)
        if self._Z.numel() != gs * gs or self.costmap_wp.tensor.numel() != cms * cms:
            raise capi.MppiError("terrain arrays do not match surface.grid_size / costmap_size")
        t = capi.MppiTerrain(self._Z.data_ptr(), gs, float(s.half_width), float(s.resolution),
                             self.costmap_wp.tensor.data_ptr(), cms, float(s.costmap_resolution))
        capi.check(self._lib.mppi_set_terrain(self._handle, C.byref(t)), "mppi_set_terrain")
        self._terrain_dirty = False

    # ------------------------------------------------------------------ per-step
    def reset(self, controller_or_sim):
        """The reference rebuilds per-sample scratch here (MPPI_isaac.py:489-503); the fused kernel keeps all
        per-sample state in registers, so there is nothing to reset."""
        return None

    def _state(self) -> capi.MppiState:
        r = self.robot
        hv = np.asarray(r.heading_vector)                                # dtype kept: run() stores a float32 array
        hv = hv / np.linalg.norm(hv)                                     # MPPI_isaac.py:493
        return capi.MppiState(float(r.x[-1]), float(r.y[-1]), float(hv[0]), float(hv[1]), float(hv[2]),
                              float(r.left_wheel_speed), float(r.right_wheel_speed),
                              float(self.std_dev_u1), float(self.std_dev_u2),
                              float(self.goal_x), float(self.goal_y), float(self.goal_orientation))

    def inject_noise(self, eps: Optional[torch.Tensor]):
        """Validation mode: eps is a device float32 tensor [2, K, T] of standard-normal noise used by the next
        steps instead of the Philox stream (None restores production mode)."""
        if eps is not None:
            K, T = self.number_of_trajectories, self.number_of_iterations
            if tuple(eps.shape[-3:]) != (2, K, T) or eps.dim() not in (3, 4) or eps.dtype != torch.float32 \
                    or not eps.is_cuda:
                raise ValueError("eps must be a float32 CUDA tensor of shape [2, K, T] ([n, 2, K, T] for device_loop)")
            eps = eps.contiguous()
        self._injected_noise = eps

    def MPPI_step(self, proj="3d"):
        """One MPPI iteration (replaces MPPI_isaac.py:505-720).  Asynchronous; results stay on the device."""
        if self._handle is None:
            raise capi.MppiError("call warp_setup() first")
        if self._terrain_dirty:
            self._push_terrain()
        pj = capi.PROJ_3D if proj == "3d" else capi.PROJ_2D
        st = self._state()
        self._stream = torch.cuda.current_stream(self.device).cuda_stream
        noise = self._injected_noise.data_ptr() if self._injected_noise is not None else None
        offset = self._step_count
        capi.check(self._lib.mppi_step(self._handle, C.byref(st), pj, noise, self.seed, offset, self._stream),
                   "mppi_step")
        self._last_state, self._last_proj, self._last_offset = st, pj, offset
        self._step_count += 1
        self._sim_stale = True

    def step_command(self, proj="3d"):
        """MPPI_step + read-back of the command (v*[0], w*[0]) in one library call (8-byte pinned D2H):
        what the driver does with MPPI_step and two `.numpy()[0]` reads (driver :468-472)."""
        if self._terrain_dirty:
            self._push_terrain()
        pj = capi.PROJ_3D if proj == "3d" else capi.PROJ_2D
        st = self._state()
        offset = self._step_count
        self._stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self._lib.mppi_step_host(self._handle, C.byref(st), pj, self.seed, offset, self._cmd,
                                            self._stream), "mppi_step_host")
        self._last_state, self._last_proj, self._last_offset = st, pj, offset
        self._step_count += 1
        self._sim_stale = True
        return float(self._cmd[0]), float(self._cmd[1])

    def _ensure_sim(self):
        """Launch 9 of the reference (optimal-trajectory rollout, MPPI_isaac.py:696-720), computed lazily."""
        if self._sim_stale and self._last_state is not None:
            capi.check(self._lib.mppi_sim_rollout(self._handle, C.byref(self._last_state), self._stream),
                       "mppi_sim_rollout")
            self._sim_stale = False

    def stats(self) -> dict:
        s = self._stats.cpu().numpy()
        i = s.view(np.int32)
        return dict(min_cost=float(s[0]), argmin=int(i[1]), weights_sum=float(s[2]), oob=int(i[3]), nan=int(i[4]),
                    ess=float(s[5]), v0=float(s[6]), w0=float(s[7]))

    def debug_dump(self, which=("traj",), replay_last: bool = True) -> dict:
        """Materialises K x T intermediates of the LAST step on request (the visualiser's `trajectories`,
        driver :520-528).  Returns torch tensors on the device."""
        K, T = self.number_of_trajectories, self.number_of_iterations
        dev = self.device
        shapes = {"u1": (K, T), "u2": (K, T), "v": (K, T), "w": (K, T), "traj": (K, T, 3), "heading": (K, T, 3),
                  "lw": (K, T, 3), "rw": (K, T, 3), "dem_ij": (K, T, 2), "lw_ij": (K, T, 2), "rw_ij": (K, T, 2),
                  "cm_ij": (K, T, 2), "critics": (K, 4), "weights": (K,), "critics_ext": (K, 6)}
        d = capi.MppiDebugDump()
        out = {}
        for name in which:
            dt = torch.int32 if name.endswith("_ij") else torch.float32
            out[name] = torch.zeros(shapes[name], dtype=dt, device=dev)
            setattr(d, name, out[name].data_ptr())
        st = self._last_state if (replay_last and self._last_state is not None) else self._state()
        noise = self._injected_noise.data_ptr() if self._injected_noise is not None else None
        capi.check(self._lib.mppi_debug_dump(self._handle, C.byref(st), self._last_proj, noise, self.seed,
                                             self._last_offset, 1 if replay_last else 0, C.byref(d), self._stream),
                   "mppi_debug_dump")
        return out

    def visualiser_points(self, block_x_current: float, block_y_current: float, half_block: float,
                          sample_stride: int = 50, step_stride: int = 10):
        """What the driver hands to VisualizeMPPI (visual_terrain_stack_full_terrain.py:252-261, 520-528):
        every `sample_stride`-th sampled trajectory of the last step at every `step_stride`-th step, moved into the
        world frame (x_w = -y + block_x + half_block, y_w = x + block_y + half_block), and the matching costs
        normalised as `(c - min(c)) / max(c)` and repeated per point.  Returns NumPy (points [n, 3], costs [n])."""
        K, T = self.number_of_trajectories, self.number_of_iterations
        ne, nt = -(-K // sample_stride), -(-T // step_stride)
        pts = torch.empty((ne, nt, 3), dtype=torch.float32, device=self.device)
        st = self._last_state if self._last_state is not None else self._state()
        noise = self._injected_noise.data_ptr() if self._injected_noise is not None else None
        capi.check(self._lib.mppi_export_trajectories(self._handle, C.byref(st), self._last_proj, noise, self.seed,
                                                      self._last_offset, 1, sample_stride, step_stride,
                                                      pts.data_ptr(), self._stream), "mppi_export_trajectories")
        world = self.to_world_frame(pts.reshape(-1, 3), block_x_current, block_y_current, half_block)
        costs = self.normalised_costs(self.costs_wp.tensor[::sample_stride])
        return world.cpu().numpy(), costs.repeat_interleave(nt).cpu().numpy()

    @staticmethod
    def to_world_frame(p: torch.Tensor, block_x_current: float, block_y_current: float, half_block: float):
        """Block frame -> world frame of the driver's `transform_trajs` (visual_terrain_stack_full_terrain.py:257-259):
        float32 additions in its order, (-y + block_x) + half_block and (x + block_y) + half_block; z unchanged."""
        return torch.stack([(-p[:, 1] + block_x_current) + half_block, (p[:, 0] + block_y_current) + half_block,
                            p[:, 2]], dim=1)

    @staticmethod
    def normalised_costs(costs: torch.Tensor):
        """`(costs - min(costs)) / max(costs)` of the driver's visualiser feed (:524)."""
        return (costs - costs.min()) / costs.max()

    def on_block_change(self, shift_x: float, shift_y: float, dem, rocks_data, origin):
        """The driver's terrain-block change (visual_terrain_stack_full_terrain.py:546-576) in one call: rebuild the
        obstacle costmap for the new block on the device, swap the DEM pointer (zero-copy when `dem` is a device
        array), and shift the robot history and the goal into the new block frame."""
        self.rebuild_costmap(rocks_data, origin)
        self.Z_wp = dem
        for j in range(len(self.robot.x)):
            self.robot.x[j] -= shift_y
            self.robot.y[j] += shift_x
        self.goal_x -= shift_y
        self.goal_y += shift_x
        self.goal = (self.goal_x, self.goal_y)

    @property
    def trajectories(self):
        """K*T x 3 sampled trajectories of the last step (MPPI_isaac.py:467), materialised on access."""
        K, T = self.number_of_trajectories, self.number_of_iterations
        return DeviceArray(self.debug_dump(("traj",))["traj"].reshape(K * T, 3))

    # ------------------------------------------------------------------ offline closed loop
    def run(self, proj="3d", max_loops: int = 3500, device_loop: bool = False):
        """Closed loop with the controller's own model as the plant, MPPI_isaac.py:755-805.

        device_loop=False walks the reference's host loop (one MPPI_step + read-backs per iteration).
        device_loop=True runs the same loop resident on the device (mppi_run_closed_loop: one fused launch per
        iteration, pose / sigma / wheel-speed feedback and the goal test inside the kernel) and then fills the
        robot's history exactly as the host loop would have."""
        self.warp_setup()
        if device_loop:
            return self._run_on_device(proj, max_loops)
        while (abs(self.robot.x[-1] - self.goal_x) > 0.5 or abs(self.robot.y[-1] - self.goal_y) > 0.5) \
                and self.loop < max_loops:
            self.reset("controller")
            self.MPPI_step(proj=proj)
            traj0 = self.trajectories_sim.numpy()[0]
            head0 = self.heading_vectors_sim.numpy()[0]
            self.robot.update_position(traj0[0], traj0[1], traj0[2], head0)
            lin_vel = self.optimal_lin_vel_wp.numpy()[0]
            ang_vel = self.optimal_ang_vel_wp.numpy()[0]
            self.std_dev_u1 = np.maximum(0.4, 0.4 - ang_vel * ang_vel)          # MPPI_isaac.py:777-778
            self.std_dev_u2 = np.maximum(0.4, 0.4 + ang_vel * ang_vel)
            self.robot.lin_vel.append(lin_vel)
            self.robot.ang_vel.append(ang_vel)
            self.robot.left_wheel_speed = lin_vel - ang_vel * self.robot.radius / 2
            self.robot.right_wheel_speed = lin_vel + ang_vel * self.robot.radius / 2
            self.loop += 1
        return self.loop

    def _run_on_device(self, proj, max_loops):
        if self._terrain_dirty:
            self._push_terrain()
        n = max_loops - self.loop
        if n <= 0 or not (abs(self.robot.x[-1] - self.goal_x) > 0.5 or abs(self.robot.y[-1] - self.goal_y) > 0.5):
            return self.loop
        pj = capi.PROJ_3D if proj == "3d" else capi.PROJ_2D
        st = self._state()
        log = np.zeros((n, 8), np.float32)
        done, reached = C.c_int32(0), C.c_int32(0)
        noise = self._injected_noise
        if noise is not None and noise.dim() == 3:
            raise ValueError("device_loop needs injected noise of shape [iterations, 2, K, T]")
        self._stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self._lib.mppi_run_closed_loop(self._handle, C.byref(st), pj,
                                                  noise.data_ptr() if noise is not None else None, self.seed,
                                                  self._step_count, n, 0.5, 0.4, 1.0, log.ctypes.data,
                                                  C.byref(done), C.byref(reached), self._stream),
                   "mppi_run_closed_loop")
        k = int(done.value)
        for row in log[:k]:
            self.robot.update_position(row[0], row[1], row[2], row[3:6].copy())
            self.robot.lin_vel.append(row[6])
            self.robot.ang_vel.append(row[7])
        self.std_dev_u1, self.std_dev_u2 = np.float32(st.sigma1), np.float32(st.sigma2)
        self.robot.left_wheel_speed, self.robot.right_wheel_speed = np.float32(st.wheel_l), np.float32(st.wheel_r)
        # replays (.trajectories, debug_dump, visualiser_points) must show the fan of rollouts the LAST executed
        # iteration really sampled: its input state and its offset, not the state after its plant step
        if k > 0:
            last_in = capi.MppiState()
            capi.check(self._lib.mppi_closed_loop_last_input(self._handle, C.byref(last_in)),
                       "mppi_closed_loop_last_input")
            self._last_state, self._last_proj, self._last_offset = last_in, pj, self._step_count + k - 1
            if noise is not None:
                self._injected_noise = noise[k - 1]
        self._step_count += k
        self.loop += k
        self._sim_stale = True
        return self.loop

    def close(self):
        if self._handle is not None:
            self._lib.mppi_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
