#!/usr/bin/env python
"""bench.py -- MPPI hot-path benchmark (one JSON line on stdout, contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1..C5] [--math strict|fast] [--critics auto|reference|ext]
                  [--exchange p2p|nccl] [--variant auto|mono|pipe] [--no-extras]
  python bench.py --impl reference ...        # CPU port of the reference path on the host cores
  tools/ab.sh <other libmppi_b200.so> [bench args]   # same-node A/B of two builds (MPPI_B200_LIB selects the library)

A "step" is one full control iteration (sample -> wheel filter -> rollout on the DEM -> critics -> softmax
update -> (v*, w*)) of ONE fused kernel launch over synthetic terrain of the BASELINE.json shape.

TOP-LEVEL LINE.  N = 1 runs BASELINE config 2 (K = 4096, T = 100, 1500^2 DEM, 750^2 costmap).  N > 1 keeps that
per-GPU workload (weak scaling): one logical controller with N x 4096 samples, sample-sharded, whose only exchange is
the softmax partial per block: flag-in-data lines stored into every rank's buffer over NVLink peer memory inside the
fused launch (default), or one NCCL all-gather + combine kernel (--exchange nccl).

SAME RUN, SAME LINE (skipped with --no-extras or a non-default --workload):
  "sharded_check"  one step of the sharded controller with both transports against the UNSHARDED controller over
                   K_total samples on one GPU (same block shape): bitwise for the fused exchange, relative for NCCL,
                   argmin equal, all ranks identical; at two temperatures (0.3: argmin-like, 2000: every sample carries
                   weight).  A mismatch makes the process exit non-zero.
  "c3"             BASELINE config 3, K = 262144 samples, T = 100, 2048^2 DEM: "weak" (262144 samples per GPU) and
                   "strong" (262144 samples in total) sample sharding, each with its own equality check.
  "c4"             BASELINE config 4, rover-sharded batch (512 rovers x K = 1024 x T = 64 per GPU), no communication.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_STEP = 230.0      # SURVEY.md Appendix B (3-D skid-steer path)
GATHER_BYTES_PER_SAMPLE_STEP = 28.0   # 7 gathered fp32 words (4 DEM corners + 2 wheel cells + 1 costmap cell)
GATHER_SECTORS_PER_SAMPLE_STEP = 5.0  # 2 corner rows + 2 wheel cells + 1 costmap cell, 32-byte sectors (worst case)
STATE_BYTES = 48                  # sizeof(MppiState): the per-step host input
CMD_BYTES = 8                     # (v*[0], w*[0]) read back per step
PIPE_MAX_SAMPLES = 148 * 32 * 2   # pick_launch (mppi_capi.cu): pipelined kernel up to this many samples in flight


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--math", default="strict", choices=["strict", "fast"])
    ap.add_argument("--variant", default="auto", choices=["auto", "mono", "pipe"], help="fused-kernel variant")
    ap.add_argument("--exchange", default=os.environ.get("MPPI_SHARD_EXCHANGE", "p2p"), choices=["p2p", "nccl"],
                    help="multi-GPU exchange of the softmax partials: fused peer-memory stores or NCCL all-gather")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between timed iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the closed-loop (run()) measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip sharded_check and the c3 / c4 sub-objects")
    ap.add_argument("--latency-steps", type=int, default=1000,
                    help="back-to-back iterations of the latency leg (p50 / p99 do not depend on --steps)")
    ap.add_argument("--extras-steps", type=int, default=30, help="timed iterations of each c3 / c4 sub-measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--rovers", type=int, default=512, help="C4: rovers per GPU (4096 rovers / 8 GPUs)")
    ap.add_argument("--critics", choices=["auto", "reference", "ext"], default="auto",
                    help="reference: the four active critics of the reference; ext: + body-slope (critics_warp.py:131-166), "
                         "roll and pitch critics (MppiParams.cw_slope_path / cw_roll / cw_pitch); auto: ext for C5, the "
                         "configuration BASELINE.json names with roll/pitch/slope critics, reference otherwise")
    ap.add_argument("--dem-noise", type=float, default=0.0,
                    help="diagnostic: add Gaussian cell-to-cell noise of this sigma (m) to the synthetic DEM (slopes then "
                         "change by degrees per step and the STRICT tangent normalisation leaves its MUFU-free window)")
    ap.add_argument("--start", default="bench", choices=["bench", "rocks"],
                    help="bench: the reference's start (-60.57, -60.23), flat ground outside the rock field; rocks: a start "
                         "inside the rock field on a crater wall (the C2rock golden scenario)")
    ap.add_argument("--K", type=int, default=0, help="override the workload's samples per GPU (diagnostics)")
    ap.add_argument("--T", type=int, default=0, help="override the workload's horizon (diagnostics)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def live_peaks(device):
    """FP32 FMA throughput and L2 gather rate measured NOW on this device (csrc/peaks.cu): the roofline denominators
    MEASURED_PEAKS.json has no figure for."""
    from mppi_b200 import capi
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    if not hasattr(capi.lib(), "mppi_measure_peaks"):        # A/B against an older library (MPPI_B200_LIB)
        clk = peaks()["sm_max_mhz"]
        return dict(fp32_tflops=148 * 128 * 2 * clk * 1e6 / 1e12, l2_gather_gsectors=float("nan"), l2_gather_gbs=float("nan"),
                    how="NOT measured (library without mppi_measure_peaks): 148 SM x 128 lanes x 2 x sm_max_mhz")
    capi.check(capi.lib().mppi_measure_peaks(device, 8 << 20, C.byref(a), C.byref(b), C.byref(c)), "mppi_measure_peaks")
    return dict(fp32_tflops=float(a.value), l2_gather_gsectors=float(b.value), l2_gather_gbs=float(c.value),
                how="csrc/peaks.cu on this device just before the timed region: 8 independent FFMA chains per thread, "
                    "8 x 256 threads per SM; random 4-byte .cg gathers in an 8 MiB L2-resident window")


def kernel_counters(workload, math, K, T, variant):
    """Per-launch counters of the dominant kernel from the committed ncu capture (profiles/kernel_counters.json).
    They describe ONE configuration: used as they are only when K, T and the kernel match; scaled by K T and marked
    extrapolated for another size of the same kernel; absent otherwise."""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        c = json.load(f).get(f"{workload}_{math}")
    if not c or c.get("variant", variant) != variant:
        return {}
    K0, T0 = c.get("K"), c.get("T")
    if K0 is None or T0 is None:
        return {}
    if (K0, T0) == (K, T):
        return dict(c, extrapolated=False)
    s = (K * T) / float(K0 * T0)
    return dict(c, warp_inst_per_launch=c["warp_inst_per_launch"] * s, dram_bytes_per_launch=None, extrapolated=True)


ROCK_START = dict(start=(-16.37, -31.73), heading=(0.6, -0.8, 0.0), goal=(20.0, -48.0))    # the C1rock / C2rock golden scene


def build_workload(name, K_override=0, T_override=0, dem_noise=0.0, start_kind="bench"):
    from mppi_b200 import synthetic as syn
    import dataclasses
    w = syn.WORKLOADS[name]
    if dem_noise > 0.0:
        w = dataclasses.replace(w, name=f"{w.name} [DEM + N(0, {dem_noise} m) per cell]")
    if K_override or T_override:
        w = dataclasses.replace(w, K=K_override or w.K, T=T_override or w.T,
                                name=f"{w.name} [override K={K_override or w.K} T={T_override or w.T}]")
    dem = syn.crater_dem(w.grid_size, w.half_width).numpy()
    if dem_noise > 0.0:
        dem = dem + np.random.default_rng(2024).normal(0.0, dem_noise, dem.shape).astype(np.float32)
    cm = syn.rock_costmap(w.costmap_size, w.half_width)
    start, goal = syn.workload_start_goal(w)
    heading = (1.0, 0.0, 0.0)
    if start_kind == "rocks":
        s = w.scale
        start = (ROCK_START["start"][0] * s, ROCK_START["start"][1] * s)
        goal = (ROCK_START["goal"][0] * s, ROCK_START["goal"][1] * s)
        heading = ROCK_START["heading"]
        w = dataclasses.replace(w, name=f"{w.name} [start inside the rock field]")
    return w, dem, cm, start, goal, heading


EXT_CRITICS = dict(cw_slope_path=50.5, cw_roll=400.0, cw_pitch=250.0)


def critic_weights(args):
    """Optional critic weights of this run (see --critics)."""
    ext = args.critics == "ext" or (args.critics == "auto" and args.workload in ("C5", "C5many"))
    return dict(EXT_CRITICS) if ext else {}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(w, dem, cm, start, goal, seconds, steps=None, warmup=0, K=None, critics=None, heading=(1.0, 0.0, 0.0)):
    """Times the C port of the reference path (oracle/mppi_oracle.c, libm math) on all host cores."""
    from oracle import oracle_c as oc
    oc.build()
    K = K or w.K
    T = w.T
    threads = oc.num_threads()
    p = oc.make_params(K=K, T=T, math=oc.MATH_LIBM, **(critics or {}))
    hn = np.asarray(heading, np.float64) / np.linalg.norm(heading)
    st = dict(x=start[0], y=start[1], hx=hn[0], hy=hn[1], hz=hn[2], wheel_l=0.0, wheel_r=0.0, sigma1=0.25, sigma2=0.25,
              goal_x=goal[0], goal_y=goal[1], goal_theta=2.2)
    rng = np.random.default_rng(0)
    e1 = rng.standard_normal((K, T)).astype(np.float32)
    e2 = rng.standard_normal((K, T)).astype(np.float32)
    n1 = np.zeros(T, np.float32)
    n2 = np.zeros(T, np.float32)
    for _ in range(max(1, warmup)):
        r = oc.mppi_step(p, dem, w.half_width, cm, st, n1, n2, e1, e2, nthreads=threads)
    times = []
    t_end = time.perf_counter() + (seconds or 0.0)
    while (steps is None and time.perf_counter() < t_end) or (steps is not None and len(times) < steps):
        t0 = time.perf_counter()
        r = oc.mppi_step(p, dem, w.half_width, cm, st, n1, n2, e1, e2, nthreads=threads)
        times.append(time.perf_counter() - t0)
        n1, n2 = r.nominal1, r.nominal2          # closed-loop dependence on the nominal, as on the GPU
    times = np.array(times)
    return dict(value=K * T / times.mean(), unit="sample-steps/s", cores=threads, kind="port",
                sample=f"{len(times)} full iterations of K={K} T={T} (injected noise, sampling excluded), "
                       f"C port of the reference kernels (oracle/mppi_oracle.c, libm), {threads} pthreads",
                ms_per_step=float(times.mean() * 1e3), p50_ms=float(np.median(times) * 1e3)), times


def cpu_reference_script_run(w, samples=32):
    """BASELINE.md baseline B0: the reference's OWN CPU projection function `generate_trajectory_25D`
    (thesis_master/python_mppi_projection/displacement_on_surface.py:317-369, extracted as written into oracle/_ref/ by
    oracle/make_ref.py) timed on this machine, one trajectory at a time on one core, as the script runs it.  Rollout
    only: the script has no wheels, critics or update, so this is a floor of the reference's CPU cost of an iteration."""
    from oracle import make_ref
    ns = make_ref.load()
    if ns is None:
        return {"unavailable": "oracle/_ref/displacement_functions.py not built (python oracle/make_ref.py in the build container)"}
    from mppi_b200 import synthetic as syn
    T = w.T
    X, Y, Z = ns["create_surface"](w.grid_size, w.half_width, syn.NINE_CRATERS[:3])
    res = 2 * w.half_width / w.grid_size
    rng = np.random.default_rng(0)
    times = []
    for k in range(samples + 2):
        v = rng.uniform(0.2, 2.0, T)
        om = rng.uniform(-1.0, 1.0, T)
        t0 = time.perf_counter()
        ns["generate_trajectory_25D"](-10.0, -10.0, np.array([1.0, 0.0, 0.0]), v, om, 0.045, T, res, X, Y, Z)
        if k >= 2:
            times.append(time.perf_counter() - t0)
    per = float(np.median(times)) / T
    return {"value": 1.0 / per, "unit": "sample-steps/s", "cores": 1, "kind": "reference",
            "us_per_sample_step": per * 1e6, "iteration_rollout_only_s": per * w.K * T,
            "sample": f"{samples} calls of the reference's generate_trajectory_25D (displacement_on_surface.py:317-369 as "
                      f"written, via oracle/_ref), T={T}, one core, rollout only (no wheels / critics / update)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, dem, cm, start, goal, heading = build_workload(args.workload, args.K, args.T, args.dem_noise, args.start)
    # like for like with the repo arm at N GPUs: ONE logical controller over K_total = N x K samples (C4: N x rovers)
    n = max(1, args.gpus)
    if args.workload == "C4":
        K_total, units_note = w.K, f"{args.rovers * n} rovers of K={w.K} T={w.T}: ONE rover iteration timed per step, value scaled by 1"
    else:
        K_total, units_note = w.K * n, f"K_total = {n} x {w.K}"
    base, times = cpu_reference_run(w, dem, cm, start, goal, seconds=None, steps=args.steps, warmup=max(3, min(args.warmup, 5)),
                                    K=K_total, critics=critic_weights(args), heading=heading)
    line = {
        "impl": "reference", "metric": "MPPI sample-steps/s", "value": base["value"], "unit": "sample-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "K_per_gpu": w.K, "K_total": K_total, "T": w.T,
                   "dem": f"{w.grid_size}x{w.grid_size} f32", "costmap": f"{w.costmap_size}x{w.costmap_size} f32",
                   "critics": "reference 4 + body-slope, roll, pitch" if critic_weights(args) else "reference 4",
                   "problem": units_note,
                   "note": "the reference's GPU path needs NVIDIA Warp (absent, no network); this arm is the C port "
                           "of its kernels on the host cores, on the same K_total as the repo arm at this --gpus"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "latency_us": {"p50": float(np.median(times) * 1e6), "p99": float(np.percentile(times, 99) * 1e6)},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ shared GPU plumbing
class Ctx:
    """Process-wide state of the GPU arm: rank / world, device, the L2 flush buffer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.do_flush = not args.no_flush

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_equal(self, arr: np.ndarray) -> bool:
        """True when every rank holds bitwise the same float32 array."""
        if self.world == 1:
            return True
        mine = self.torch.from_numpy(np.ascontiguousarray(arr, np.float32).view(np.int32).copy()).to(self.dev)
        every = self.torch.empty((self.world, mine.numel()), dtype=self.torch.int32, device=self.dev)
        self.dist.all_gather_into_tensor(every, mine)
        return bool((every == every[0]).all().item())

    def timed_loop(self, fn, n, first, flush):
        """n launches of fn(i), each bracketed by CUDA events on the current stream; returns ms per launch."""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            if flush:
                self.flush_buf.fill_(i & 0xff)
            evs[i][0].record()
            fn(first + i)
            evs[i][1].record()
        torch.cuda.synchronize(self.dev)
        return np.array([a.elapsed_time(b) for a, b in evs], dtype=np.float64)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def variant_id(name):
    from mppi_b200 import capi
    return {"auto": capi.VARIANT_AUTO, "mono": capi.VARIANT_MONO, "pipe": capi.VARIANT_PIPE}[name]


def make_sharded(ctx, K_local, T, dem, cm, hw, transport, math="strict", variant="auto", **overrides):
    """(core, stepper-or-None) of one logical controller with world x K_local samples on this rank's device."""
    from mppi_b200.core import Core
    from mppi_b200.sharding import SampleShardedStepper
    core = Core(K_local, T, device=ctx.local_rank, math=math, variant=variant_id(variant), **overrides)
    core.set_terrain(dem, hw, cm)
    stepper = SampleShardedStepper(core, K_local * ctx.world, transport=transport) if ctx.world > 1 else None
    return core, stepper


def read_result(core):
    ctx_sync = core.optimal_u1.device
    import torch
    torch.cuda.synchronize(ctx_sync)
    return dict(u1=core.optimal_u1[0].cpu().numpy().copy(), u2=core.optimal_u2[0].cpu().numpy().copy(),
                v=core.optimal_v[0].cpu().numpy().copy(), **core.read_stats())


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))


def equality_check(ctx, K_local, T, dem, cm, hw, state, lambdas=(0.3, 2000.0), transports=("p2p", "nccl"),
                   same_block_shape=True, math="strict"):
    """One step of the sample-sharded controller (every transport) against the UNSHARDED controller over
    K_total = world x K_local samples on this GPU, from a warm nominal, per temperature.  `same_block_shape`: the
    unsharded launch is forced onto the kernel the shards use (the pipelined one, 32 samples per block), which is what
    makes the fused exchange BITWISE comparable; otherwise the comparison is relative."""
    from mppi_b200 import capi
    from mppi_b200.core import Core
    K_total = K_local * ctx.world
    nom = np.full(T, 0.35, np.float32)
    pipe_shards = K_local <= PIPE_MAX_SAMPLES
    out = {"K_total": K_total, "transports": list(transports), "lambdas": list(lambdas), "passed": True, "cases": []}
    for lam in lambdas:
        full = Core(K_total, T, device=ctx.local_rank, math=math, lambda_=lam,
                    variant=capi.VARIANT_PIPE if (same_block_shape and pipe_shards) else capi.VARIANT_AUTO)
        full.set_terrain(dem, hw, cm)
        full.set_nominal(nom, nom)
        full.step(state, capi.PROJ_3D, None, 5, 11)
        ref = read_result(full)
        full.close()
        case = {"lambda": lam, "ess_unsharded": ref["ess"], "argmin_unsharded": ref["argmin"]}
        for tr in transports:
            if ctx.world == 1 and tr == "nccl":
                continue
            core, stepper = make_sharded(ctx, K_local, T, dem, cm, hw, tr, math=math, lambda_=lam)
            got = None
            for it in range(3):                      # three iterations: both parities of the exchange buffers
                core.set_nominal(nom, nom)
                if stepper is None:
                    core.step(state, capi.PROJ_3D, None, 5, 11)
                else:
                    stepper.step(state, capi.PROJ_3D, 5, 11)
                got = read_result(core)
            bitwise = bool(np.array_equal(got["u1"], ref["u1"]) and np.array_equal(got["u2"], ref["u2"])
                           and np.array_equal(got["v"], ref["v"]))
            r = max(rel_err(got["u1"], ref["u1"]), rel_err(got["u2"], ref["u2"]))
            same = ctx.all_equal(np.concatenate([got["u1"], got["u2"], got["v"]]))
            ok = (got["argmin"] == ref["argmin"]) and (got["min_cost"] == ref["min_cost"]) and same and r < 1e-5
            if tr == "p2p" and same_block_shape and pipe_shards:
                ok = ok and bitwise
            case[tr] = {"bitwise": bitwise, "rel": r, "argmin_equal": got["argmin"] == ref["argmin"],
                        "ranks_identical": same, "ok": bool(ok)}
            out["passed"] = out["passed"] and bool(ok)
            core.close()
        out["cases"].append(case)
    # every rank must agree on the verdict
    out["passed"] = ctx.max_over_ranks(0.0 if out["passed"] else 1.0) == 0.0
    # the summary keys the round-1 verdict asked for
    first = out["cases"][0]
    out["p2p_bitwise"] = all(c.get("p2p", {}).get("bitwise", True) for c in out["cases"])
    out["nccl_rel"] = max([c["nccl"]["rel"] for c in out["cases"] if "nccl" in c], default=None)
    out["argmin_equal"] = all(v["argmin_equal"] for c in out["cases"] for k, v in c.items() if isinstance(v, dict))
    del first
    return out


def measure_sample_sharded(ctx, core, stepper, state, K_local, T, steps, warmup, seed=42):
    """Device-timed sample-sharded iterations: `steps` launches, CUDA events, barrier + synchronise on both sides,
    max over ranks.  Returns (ms_per_step, per-step ms of this rank, value in sample-steps/s)."""
    from mppi_b200 import capi

    def one(i):
        if stepper is None:
            core.step(state, capi.PROJ_3D, None, seed, i)
        else:
            stepper.step(state, capi.PROJ_3D, seed, i)

    ctx.timed_loop(one, max(3, warmup), 0, ctx.do_flush)
    ctx.barrier()
    per = ctx.timed_loop(one, steps, 1000, ctx.do_flush)
    ctx.barrier()
    ms = ctx.max_over_ranks(float(per.sum())) / steps
    return ms, per, K_local * ctx.world * T / (ms * 1e-3), one


def run_c3(ctx, args):
    """BASELINE config 3 inside the default run: weak (K = 262144 per GPU) and strong (K = 262144 in total)."""
    import torch
    from mppi_b200 import synthetic as syn
    from mppi_b200.core import make_state
    w = syn.WORKLOADS["C3"]
    dem = syn.crater_dem(w.grid_size, w.half_width, device=ctx.dev).contiguous()
    cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(ctx.dev)
    start, goal = syn.workload_start_goal(w)
    state = make_state(start[0], start[1], (1.0, 0.0, 0.0), goal_x=goal[0], goal_y=goal[1])
    out = {"workload": w.name, "T": w.T, "dem": f"{w.grid_size}x{w.grid_size} f32", "math": args.math,
           "exchange": args.exchange, "steps": args.extras_steps}
    for mode, K_local in (("weak", w.K), ("strong", w.K // ctx.world)):
        if mode == "strong" and ctx.world == 1:
            out["strong"] = dict(out["weak"], note="N = 1: identical to weak")
            continue
        core, stepper = make_sharded(ctx, K_local, w.T, dem, cm, w.half_width, args.exchange, math=args.math)
        ms, per, value, _ = measure_sample_sharded(ctx, core, stepper, state, K_local, w.T, args.extras_steps, 5)
        core.close()
        chk = equality_check(ctx, K_local, w.T, dem, cm, w.half_width, state, lambdas=(0.3,), transports=(args.exchange,),
                             same_block_shape=False, math=args.math)
        out[mode] = {"K_per_gpu": K_local, "K_total": K_local * ctx.world, "ms_per_step": ms, "value": value,
                     "unit": "sample-steps/s", "p50_us": float(np.median(per) * 1e3),
                     "check": {"passed": chk["passed"], "rel": chk["cases"][0][args.exchange]["rel"],
                               "argmin_equal": chk["argmin_equal"],
                               "ranks_identical": chk["cases"][0][args.exchange]["ranks_identical"]}}
    return out


def many_start_states(w, n_side=8):
    """C5many: n_side^2 poses on a grid spread over the map, headings fanned out, goals mirrored through the centre."""
    from mppi_b200.core import make_state
    span = w.half_width * 0.75
    xs = np.linspace(-span, span, n_side)
    states = []
    for j, y in enumerate(xs):
        for i, x in enumerate(xs):
            a = 2 * np.pi * ((i * n_side + j) % 16) / 16.0
            states.append(make_state(float(x), float(y), (float(np.cos(a)), float(np.sin(a)), 0.0),
                                     goal_x=float(-x), goal_y=float(-y)))
    return states


def run_c4(ctx, args, K=0, T=0, many=False):
    """BASELINE config 4: independent controllers sharded by rover, no communication.  Each GPU runs `--rovers`
    rovers (default 512 = 4096 / 8) x K = 1024 x T = 64 in ONE launch (grid.y = rover); every rover has its own DEM /
    costmap memory, pose, goal, nominal and Philox stream.  Returns the measurement as a dict.

    many=True (--workload C5many): BASELINE config 5's 8192^2 DEM (268 MB) with 64 controllers x K = 1024 x T = 200
    started on an 8 x 8 grid of poses over ONE shared map: the touched windows add up to ~530 MB, so -- unlike the
    single-start C5, whose 8 MB window stays in L2 -- the terrain gathers really come from HBM."""
    import torch
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    w = syn.WORKLOADS["C5" if many else "C4"]
    if many:
        states = many_start_states(w)
        K, T, R = K or 1024, T or w.T, len(states)
        pool = 1
        dem = syn.crater_dem(w.grid_size, w.half_width, device=dev).contiguous()
        cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
        dems, cms = dem.unsqueeze(0).expand(R, -1, -1), cm.unsqueeze(0).expand(R, -1, -1)
        core = Core(K, T, device=ctx.local_rank, math=args.math, max_rovers=R, **critic_weights(args))
        core.set_terrain_batched_shared(dems, w.half_width, cms)
    else:
        K, T, R = K or w.K, T or w.T, args.rovers
        pool = 16                                    # distinct synthetic maps; every rover gets its OWN copy in HBM
        rng = np.random.default_rng(7 + rank)
        dem_pool = torch.stack([syn.crater_dem(w.grid_size, w.half_width, seed=57 + i, device=dev) for i in range(pool)])
        cm_pool = torch.stack([torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width, n_rocks=90, seed=99 + i))
                               for i in range(pool)]).to(dev)
        idx = torch.arange(R, device=dev) % pool
        dems, cms = dem_pool[idx].contiguous(), cm_pool[idx].contiguous()
        core = Core(K, T, device=ctx.local_rank, math=args.math, max_rovers=R)
        core.set_terrain_batched(dems, w.half_width, cms)
        half = w.half_width / 2
        states = [make_state(float(rng.uniform(-half, half)), float(rng.uniform(-half, half)),
                             (float(np.cos(a)), float(np.sin(a)), 0.0), goal_x=float(rng.uniform(-half, half)),
                             goal_y=float(rng.uniform(-half, half)))
                  for a in rng.uniform(0, 2 * np.pi, R)]
    states_host = torch.frombuffer(bytearray(bytes((capi.MppiState * R)(*states))), dtype=torch.uint8).pin_memory()
    states_dev = states_host.to(dev)
    cmd_host = torch.empty((R, 2), dtype=torch.float32).pin_memory()

    def one(i):
        core.step_batched(states_dev, R, capi.PROJ_3D, 42, i)

    steps = min(args.steps if args.workload in ("C4", "C5many") else args.extras_steps, 100)
    warm = max(3, min(args.warmup, 10))
    ctx.timed_loop(one, warm, 0, ctx.do_flush)
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ctx.barrier()
    per = ctx.timed_loop(one, steps, 1000, ctx.do_flush)
    ctx.barrier()
    clocks = sampler.stop()
    ms = ctx.max_over_ranks(float(per.sum())) / steps
    # end to end: rover states from pinned host memory in, all commands back on the host, inside the timed region
    e2e_t = []
    for i in range(min(steps, 30)):
        if ctx.do_flush:
            ctx.flush_buf.fill_(i & 0xff)
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        states_dev.copy_(states_host, non_blocking=True)
        core.step_batched(states_dev, R, capi.PROJ_3D, 42, 5000 + i)
        cmd_host.copy_(core.stats[:R, 6:8], non_blocking=True)
        torch.cuda.synchronize(dev)
        e2e_t.append(time.perf_counter() - t0)
    e2e_mean = ctx.max_over_ranks(float(np.mean(e2e_t)))
    # check: rover 0 of the batch against a stand-alone controller on rover 0's maps (same Philox stream: rover index
    # 0): per-sample costs bit for bit, update within the summation-order tolerance (the two launches use different
    # block shapes)
    core.set_nominal(np.zeros((R, T), np.float32), np.zeros((R, T), np.float32), n_rovers=R)
    core.step_batched(states_dev, R, capi.PROJ_3D, 42, 77)
    torch.cuda.synchronize(dev)
    costs_b = core.costs[0].cpu().numpy().copy()
    u1_b = core.optimal_u1[0].cpu().numpy().copy()
    solo = Core(K, T, device=ctx.local_rank, math=args.math, **(critic_weights(args) if many else {}))
    solo.set_terrain(dems[0].contiguous() if not many else dem, w.half_width, cms[0].contiguous() if not many else cm)
    solo.step(states[0], capi.PROJ_3D, None, 42, 77)
    torch.cuda.synchronize(dev)
    check = {"rover0_costs_bitwise": bool(np.array_equal(costs_b, solo.costs[0].cpu().numpy())),
             "rover0_update_rel": rel_err(u1_b, solo.optimal_u1[0].cpu().numpy()),
             "all_commands_finite": bool(torch.isfinite(core.stats[:R, 6:8]).all().item())}
    check["passed"] = bool(check["rover0_costs_bitwise"] and check["rover0_update_rel"] < 1e-5 and check["all_commands_finite"])
    check["passed"] = ctx.max_over_ranks(0.0 if check["passed"] else 1.0) == 0.0
    solo.close()
    units = R * K * T
    dur = ms * 1e-3
    res = {"workload": (w.name + " [64 starts x K=1024 on ONE shared DEM]") if many else w.name, "rovers_per_gpu": R, "rovers_total": R * world, "K": K, "T": T, "steps": steps, "warmup": warm,
           "ms_per_step": ms, "value": units * world / dur, "unit": "sample-steps/s", "rover_updates_per_s": R * world / dur,
           "p50_us": float(np.median(per) * 1e3), "p99_us": float(np.percentile(per, 99) * 1e3),
           "dem": (f"{w.grid_size}x{w.grid_size} f32, ONE map shared by all starts" if many else
                   f"{w.grid_size}x{w.grid_size} f32 per rover ({pool} distinct maps, one copy per rover)"),
           "costmap": f"{w.costmap_size}x{w.costmap_size} f32 per rover", "math": args.math,
           "parallelism": f"rover-sharded x{world}, no communication",
           "e2e": {"value": units * world / e2e_mean, "unit": "sample-steps/s", "h2d_bytes_per_step": R * STATE_BYTES,
                   "d2h_bytes_per_step": R * CMD_BYTES, "p50_us": float(np.median(e2e_t) * 1e6),
                   "call": "pinned states H2D + mppi_step_batched + all (v*, w*) D2H + stream synchronise"},
           "check": check, "clocks": clocks, "stats_rover0": core.read_stats(0)}
    core.close()
    return res


def run_rover_batch_line(ctx, args):
    """--workload C4 / C5many: the rover-sharded batch as the top-level line."""
    many = args.workload == "C5many"
    r = run_c4(ctx, args, args.K, args.T, many=many)
    if ctx.rank == 0:
        pk = peaks()
        lp = live_peaks(ctx.local_rank)
        units = r["rovers_per_gpu"] * r["K"] * r["T"]
        dur = r["ms_per_step"] * 1e-3
        line = {
            "metric": "MPPI sample-steps/s", "value": r["value"], "unit": "sample-steps/s", "n_gpus": ctx.world,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {k: r[k] for k in ("workload", "rovers_per_gpu", "rovers_total", "K", "T", "dem", "costmap", "math",
                                         "parallelism")} |
                      {"l2": "flushed between timed iterations (256 MiB fill)" if ctx.do_flush else "warm"},
            "latency_us": {"p50": r["p50_us"], "p99": r["p99_us"]},
            "rover_updates_per_s": r["rover_updates_per_s"], "e2e": r["e2e"], "gpu_launches": r["steps"],
            "roofline": roofline_object("mono", units, r["T"], dur, pk, lp,
                                        kernel_counters("C5many" if many else "C4", args.math, r["K"], r["T"], "mono")
                                        if ctx.world == 1 and r["rovers_per_gpu"] == (64 if many else 512) else {},
                                        None, "mppi_fused_kernel<3D, Philox>, grid.y = rover"),
            "check": r["check"], "clocks": r["clocks"], "stats_rover0": r["stats_rover0"],
        }
        emit(line)
        if not r["check"]["passed"]:
            sys.exit(3)


def roofline_object(kernel_kind, units, T, dur_s, pk, lp, counters, share, kernel_name):
    """The contract's roofline object for the dominant kernel.  `units` = sample-steps ONE launch processes on ONE GPU.
    The path streams nothing through HBM (no K x T tensor exists) and its gathers are served by shared memory / L1 / L2,
    so neither HBM nor the tensor cores bind it: `bound` names what does -- the dependent-issue LATENCY of one warp's
    chain in the pipelined kernel (K of a few thousand), instruction ISSUE in the monolithic kernel (large K) -- and
    the HBM / FP32 / issue / L2-gather fractions are sub-objects, each against a peak measured on this pool."""
    hbm_ach = GATHER_BYTES_PER_SAMPLE_STEP * units / dur_s / 1e9
    fp32_ach = FLOP_PER_SAMPLE_STEP * units / dur_s / 1e12
    sect_ach = GATHER_SECTORS_PER_SAMPLE_STEP * units / dur_s / 1e9
    issue_peak = 148 * 4 * pk["sm_max_mhz"] * 1e6
    r = {
        "bound": "latency" if kernel_kind == "pipe" else "issue",
        "achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"],
        "traffic": counters.get("dram_bytes_per_launch"),
        "kernel": kernel_name, "kernel_share_of_step": share,
        "per_gpu": True, "sample_steps_per_launch": units,
        "algorithmic_bytes_per_sample_step": GATHER_BYTES_PER_SAMPLE_STEP,
        "peak_source": f"{pk['source']} MEASURED_PEAKS.json hbm_gbs",
        "note": "achieved / peak / frac: algorithmic gather bytes against the measured HBM peak, as the contract asks; the "
                "gathers are served on-chip (traffic = DRAM bytes per launch from ncu when the captured configuration "
                "matches, else null), so this fraction is small by construction and `bound` is not 'hbm'",
        "hbm": {"achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"]},
        "fp32": {"achieved": fp32_ach, "peak": lp["fp32_tflops"], "unit": "TFLOP/s", "frac": fp32_ach / lp["fp32_tflops"],
                 "flop_per_sample_step": FLOP_PER_SAMPLE_STEP, "peak_source": "measured live: " + lp["how"]},
        "l2_gather": {"algorithmic": sect_ach, "peak": lp["l2_gather_gsectors"], "unit": "Gsector/s",
                      "ratio_if_nothing_were_staged": sect_ach / lp["l2_gather_gsectors"],
                      "sectors_per_sample_step": GATHER_SECTORS_PER_SAMPLE_STEP,
                      "peak_source": "measured live (random 32-byte-sector gathers from L2)",
                      "note": "NOT an achieved fraction: the worst-case sector count of the lookups (5 per sample-step) "
                              "over the launch time, against the L2 random-gather rate -- what the L2 would have to "
                              "deliver if shared memory / L1 staged nothing; above 1 means the path only works BECAUSE "
                              "they do (measured L2 traffic: l2_measured)"},
    }
    if counters.get("dram_bytes_per_launch"):
        t = counters["dram_bytes_per_launch"] / dur_s / 1e9
        r["hbm_measured"] = {"achieved": t, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": t / pk["hbm_gbs"],
                             "note": "DRAM bytes per launch from the committed ncu capture of THIS configuration "
                                     "(dram__bytes_read + write) over the launch time measured here"}
    if counters.get("l2_sectors_per_launch") and not counters.get("extrapolated"):
        g = counters["l2_sectors_per_launch"] * 32.0 / dur_s / 1e9
        r["l2_measured"] = {"achieved": g, "peak": lp["l2_gather_gbs"], "unit": "GB/s", "frac": g / lp["l2_gather_gbs"],
                            "hit_rate_pct": counters.get("l2_hit_pct"),
                            "note": "lts__t_sectors x 32 B per launch (ncu capture of this configuration) over the launch "
                                    "time measured here, against the live L2 random-gather rate"}
    if kernel_kind == "pipe":
        r["latency"] = {"ns_per_horizon_step": dur_s * 1e9 / T, "unit": "ns",
                        "note": "the whole launch divided by T: the chain warp's dependent issue latency per horizon step "
                                "plus the launch's fixed parts (tools/timeline.py splits them)"}
    if counters.get("warp_inst_per_launch"):
        ach = counters["warp_inst_per_launch"] / dur_s
        r["issue"] = {"achieved": ach, "peak": issue_peak, "unit": "warp-inst/s", "frac": ach / issue_peak,
                      "extrapolated": bool(counters.get("extrapolated")), "source": counters.get("source"),
                      "peak_source": "148 SM x 4 schedulers x sm_max_mhz"}
    return r


# ------------------------------------------------------------------------------------------ GPU arm
_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout.  Everything else that native libraries print there
    (NCCL writes its version banner to stdout) is diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    from mppi_b200 import capi
    from mppi_b200.core import make_state

    ctx = Ctx(args)
    if args.workload in ("C4", "C5many"):
        run_rover_batch_line(ctx, args)
        ctx.close()
        return
    dev, rank, n_gpus = ctx.dev, ctx.rank, ctx.world

    w, dem_np, cm_np, start, goal, heading = build_workload(args.workload, args.K, args.T, args.dem_noise, args.start)
    K, T = w.K, w.T                       # per-GPU samples
    K_total = K * n_gpus
    dem = torch.from_numpy(dem_np).to(dev)
    cm = torch.from_numpy(cm_np).to(dev)
    state = make_state(start[0], start[1], heading, goal_x=goal[0], goal_y=goal[1])
    core, stepper = make_sharded(ctx, K, T, dem, cm, w.half_width, args.exchange, math=args.math, variant=args.variant,
                                 **critic_weights(args))
    seed = 42
    lp = live_peaks(ctx.local_rank) if rank == 0 else None

    # ---- the contract's timed region: W warm-up steps, then EXACTLY K steps, barrier + synchronise on both sides
    sampler = ClockSampler(ctx.local_rank)
    warm_n = max(3, args.warmup)

    def one_step(i):
        if stepper is None:
            core.step(state, capi.PROJ_3D, None, seed, i)
        else:
            stepper.step(state, capi.PROJ_3D, seed, i)

    ctx.timed_loop(one_step, warm_n, 0, ctx.do_flush)
    sampler.start()
    ctx.barrier()
    t_wall0 = time.perf_counter()
    per_step_ms = ctx.timed_loop(one_step, args.steps, 1000, ctx.do_flush)
    ctx.barrier()
    wall_s = time.perf_counter() - t_wall0
    ms_per_step = ctx.max_over_ranks(float(per_step_ms.sum())) / args.steps
    value = K_total * T / (ms_per_step * 1e-3)

    # ---- latency leg: >= 1000 back-to-back iterations whatever --steps is, same protocol
    ctx.barrier()
    lat_ms = ctx.timed_loop(one_step, max(args.latency_steps, 1), 100000, ctx.do_flush)
    clocks = sampler.stop()
    # steady state (L2 warm) for context
    warm_ms = ctx.timed_loop(one_step, min(max(args.steps, 100), 1000), 200000, False)

    # ---- share of the step spent in the fused kernel: the library's own events around the launch vs ours
    share = lib_lat = None
    if n_gpus == 1:
        capi.check(core.L.mppi_enable_timing(core.h, 1), "mppi_enable_timing")
        ours, theirs = [], []
        us = C.c_float()
        for i in range(50):
            if ctx.do_flush:
                ctx.flush_buf.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_step(300000 + i)
            e1.record()
            capi.check(core.L.mppi_last_step_us(core.h, C.byref(us)), "mppi_last_step_us")
            torch.cuda.synchronize(dev)
            ours.append(e0.elapsed_time(e1) * 1e3)
            theirs.append(us.value)
        lib_lat = core.latency_stats() if hasattr(core.L, "mppi_latency_stats") else None
        capi.check(core.L.mppi_enable_timing(core.h, 0), "mppi_enable_timing")
        share = float(np.sum(theirs) / np.sum(ours))

    # ---- end to end through the host-facing call: host state in, host (v, w) out inside the timed region
    e2e_steps = max(args.steps, 200) if n_gpus == 1 else min(max(args.steps, 100), 300)
    if n_gpus == 1:
        for i in range(5):
            core.step_host(state, capi.PROJ_3D, seed, 9000 + i)
        e2e_t = []
        for i in range(e2e_steps):
            if ctx.do_flush:
                ctx.flush_buf.fill_(i & 0xff)
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            core.step_host(state, capi.PROJ_3D, seed, 10000 + i)
            e2e_t.append(time.perf_counter() - t0)
        e2e_t = np.array(e2e_t)
        e2e_mean = float(e2e_t.mean())
        call = ("mppi_step_host (C ABI): MppiState by value from host memory; the kernel stores the 8-byte command into "
                "mapped pinned host memory as soon as it exists, the host polls its sequence word")
    else:
        # the sharded step's command is read back on every rank; the ranks run in lock step because every step needs
        # every rank's partials, so a barrier per step would only add its own skew
        for i in range(5):
            stepper.step_host(state, capi.PROJ_3D, seed, 9000 + i)
        ctx.barrier()
        e2e_t = []
        for i in range(e2e_steps):
            if ctx.do_flush:
                ctx.flush_buf.fill_(i & 0xff)
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            stepper.step_host(state, capi.PROJ_3D, seed, 20000 + i)
            e2e_t.append(time.perf_counter() - t0)
        e2e_t = np.array(e2e_t)
        e2e_mean = ctx.max_over_ranks(float(e2e_t.mean()))
        call = "mppi_step_sharded_host on every rank (same zero-copy command store), max over ranks of the mean"
    e2e = {"value": K_total * T / e2e_mean, "unit": "sample-steps/s", "h2d_bytes_per_step": STATE_BYTES,
           "d2h_bytes_per_step": CMD_BYTES, "p50_us": float(np.median(e2e_t) * 1e6),
           "p99_us": float(np.percentile(e2e_t, 99) * 1e6), "steps": int(e2e_steps), "call": call}

    # f2: the offline closed loop of MPPI_Controller.run resident on the device (one launch per iteration, plant step
    # and feedback logic inside the kernel) vs the same loop driven from the host with a read-back per iteration.
    closed_loop = None
    if n_gpus == 1 and not args.no_closed_loop:
        import copy
        n_it = 500
        st_d = copy.copy(state)
        core.set_nominal(np.zeros(T, np.float32), np.zeros(T, np.float32))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        done, reached, _ = core.run_closed_loop(st_d, n_it, capi.PROJ_3D, seed, 50000, want_log=True)
        t_dev = time.perf_counter() - t0
        st_h = copy.copy(state)
        core.set_nominal(np.zeros(T, np.float32), np.zeros(T, np.float32))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(n_it):
            v0, w0 = core.step_host(st_h, capi.PROJ_3D, seed, 50000 + i)
            core.sim_rollout(st_h)
            pose = torch.cat([core.sim_traj[0], core.sim_heading[0]]).cpu().numpy()
            hv = pose[3:6] / np.linalg.norm(pose[3:6])
            st_h.x, st_h.y, st_h.hx, st_h.hy, st_h.hz = float(pose[0]), float(pose[1]), float(hv[0]), float(hv[1]), float(hv[2])
            w2 = np.float32(w0) * np.float32(w0)
            st_h.sigma1, st_h.sigma2 = float(max(np.float32(0.4), np.float32(0.4) - w2)), float(max(np.float32(0.4), np.float32(0.4) + w2))
            st_h.wheel_l = float(np.float32(v0) - np.float32(w0) * np.float32(1.2) / np.float32(2))
            st_h.wheel_r = float(np.float32(v0) + np.float32(w0) * np.float32(1.2) / np.float32(2))
        t_host = time.perf_counter() - t0
        closed_loop = {"iterations": int(done), "device_resident_us_per_iteration": t_dev / max(1, done) * 1e6,
                       "host_driven_us_per_iteration": t_host / n_it * 1e6,
                       "reference_published_us_per_loop": 3000.0,
                       "note": "MPPI_Controller.run (MPPI_isaac.py:755-805): plant = the controller's own model; "
                               "reference figure: 'work summarise':71 (K=1000, T=100, unspecified GPU)"}
    stats = core.read_stats()
    core.close()

    # ---- same run: equality check of the sharded controller, BASELINE configs 3 and 4
    extras = (not args.no_extras) and args.workload == "C2" and not (args.K or args.T)
    sharded_check = c3 = c4 = None
    if extras:
        sharded_check = equality_check(ctx, K, T, dem, cm, w.half_width, state, math=args.math)
        c3 = run_c3(ctx, args)
        c4 = run_c4(ctx, args)
        for k in ("clocks", "stats_rover0"):
            c4.pop(k, None)

    failed = False
    if rank == 0:
        pk = peaks()
        pipe = (K <= PIPE_MAX_SAMPLES and args.variant != "mono") or args.variant == "pipe"
        kind = "pipe" if pipe else "mono"
        counters = kernel_counters(w.name.split()[0], args.math, K, T, kind) if n_gpus == 1 else {}
        kname = ("mppi_fused_pipe_kernel<3D, Philox> (128 worker blocks + 1 updater block)" if pipe
                 else "mppi_fused_kernel<3D, Philox>")
        roofline = roofline_object(kind, K * T, T, ms_per_step * 1e-3, pk, lp, counters,
                                   share if share is not None else None, kname)
        line = {
            "metric": "MPPI sample-steps/s", "value": value, "unit": "sample-steps/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": warm_n, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "K_per_gpu": K, "K_total": K_total, "T": T,
                       "dem": f"{w.grid_size}x{w.grid_size} f32", "costmap": f"{w.costmap_size}x{w.costmap_size} f32",
                       "math": args.math, "variant": args.variant, "noise": "in-kernel Philox4x32-10 + Box-Muller", "proj": "3d",
                       "critics": ("reference 4 (path, wheel slope, speed, obstacle) + body-slope, roll, pitch"
                                   if critic_weights(args) else "reference 4 (path, wheel slope, speed, obstacle)"),
                       "scene": ("the reference's start (-60.57, -60.23) -> goal (65.8, 65.4): flat ground outside the rock "
                                 "field, no lethal cell within reach (the rock-field case: --start rocks, and the C2rock "
                                 "golden test)" if args.start == "bench" else
                                 "start inside the rock field on a crater wall (lethal cells and slopes within reach)"),
                       "l2": "flushed between timed iterations (256 MiB fill)" if ctx.do_flush else "warm",
                       "parallelism": "single GPU" if n_gpus == 1 else
                       (f"sample-sharded x{n_gpus}; softmax partial per block exchanged "
                        + ("inside the fused launch over NVLink peer memory (CUDA IPC; flag-in-data lines, no collective "
                           "call, no fence)" if args.exchange == "p2p" else "with one NCCL all-gather + combine kernel"))},
            "latency_us": {"p50": float(np.median(lat_ms) * 1e3), "p99": float(np.percentile(lat_ms, 99) * 1e3),
                           "mean": float(lat_ms.mean() * 1e3), "max": float(lat_ms.max() * 1e3), "steps": int(len(lat_ms)),
                           "steps_over_2x_p50": int((lat_ms > 2 * np.median(lat_ms)).sum()),
                           "timed_region_p50": float(np.median(per_step_ms) * 1e3),
                           "library_ring": lib_lat,
                           "note": "device time of back-to-back launches on this rank (CUDA events), L2 flushed between them; "
                                   "library_ring = mppi_latency_stats over the 50 launches of the share leg"},
            "warm_l2": {"ms_per_step": float(warm_ms.mean()), "p50_us": float(np.median(warm_ms) * 1e3),
                        "value": K_total * T / float(warm_ms.mean() * 1e-3)},
            "e2e": e2e,
            "gpu_launches": args.steps * (1 if (n_gpus == 1 or args.exchange == "p2p") else 2),
            "roofline": roofline,
            "measured_peaks_live": lp,
            "closed_loop": closed_loop,
            "clocks": clocks,
            "wall_s": wall_s,
            "stats": stats,
        }
        if extras:
            line["sharded_check"] = sharded_check
            line["c3"] = c3
            line["c4"] = c4
            failed = (not sharded_check["passed"] or not c4["check"]["passed"]
                      or any(not c3[m]["check"]["passed"] for m in ("weak", "strong") if "check" in c3[m]))
        if n_gpus == 1 and not args.no_cpu_baseline:
            base, _ = cpu_reference_run(w, dem_np, cm_np, start, goal, seconds=args.cpu_seconds,
                                        critics=critic_weights(args), heading=heading)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["reference_script"] = cpu_reference_script_run(w)
        emit(line)
    ctx.close()
    if failed:
        sys.stderr.write("bench.py: an equality check FAILED (see sharded_check / c3 / c4 in the line)\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
