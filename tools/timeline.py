#!/usr/bin/env python
"""Timeline of one fused MPPI launch from the kernel's own %globaltimer stamps (mppi_set_trace).

  python tools/timeline.py [--workload C2] [--math strict] [--variant auto] [--K 0] [--T 0] [--reps 20]

Prints, per phase, the min / median / max over blocks (microseconds relative to the first block's entry) and the
CUDA-event duration of the same launches, as one JSON object.  Diagnostic only (GPU needed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NS = 32


def sharded(a):
    """Per-rank timeline of the sharded step.  %globaltimer is per GPU, so only differences WITHIN a rank are reported:
    how long the rank's updater waits after its own last worker has published (= remote blocks that finish later +
    NVLink flight + polling), and what it costs afterwards."""
    import torch
    import torch.distributed as dist
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state
    from mppi_b200.sharding import SampleShardedStepper
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = syn.WORKLOADS[a.workload]
    if a.K or a.T:
        w = dataclasses.replace(w, K=a.K or w.K, T=a.T or w.T)
    dem = syn.crater_dem(w.grid_size, w.half_width).to(dev)
    cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
    start, goal = syn.workload_start_goal(w)
    core = Core(w.K, w.T, device=local, math=a.math)
    core.set_terrain(dem, w.half_width, cm)
    st = make_state(start[0], start[1], goal_x=goal[0], goal_y=goal[1])
    stepper = SampleShardedStepper(core, w.K * world, transport="p2p")
    nb = C.c_int32()
    capi.check(core.L.mppi_set_trace(core.h, None, C.byref(nb)), "mppi_set_trace")
    trace = torch.zeros((nb.value, NS), dtype=torch.int64, device=dev)
    capi.check(core.L.mppi_set_trace(core.h, trace.data_ptr(), None), "mppi_set_trace")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows, evt = [], []
    for rep in range(a.reps):
        dist.barrier()
        for j in range(8):
            if not a.no_flush:
                flush.fill_(j)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            stepper.step(st, capi.PROJ_3D, 42, 1000 * rep + j)
            e1.record()
        torch.cuda.synchronize()
        evt.append(e0.elapsed_time(e1) * 1e3)
        t = trace.cpu().numpy().astype(np.float64)
        rel = (t - t[:, 0].min()) / 1e3
        rel[t == 0] = np.nan
        rows.append(rel)
    rel = np.stack(rows)
    wk, upd = rel[:, :-1], rel[:, -1]
    med = lambda x: float(np.nanmedian(x))                                        # noqa: E731
    out = {"rank": rank, "world": world, "event_us_last_of_8": med(evt),
           "rollout_end_last": med(np.nanmax(wk[:, :, 3], axis=1)),
           "own_last_header_published": med(np.nanmax(wk[:, :, 5], axis=1)),
           "updater_all_headers_and_min": med(upd[:, 12]), "updater_fold": med(upd[:, 13]),
           "updater_command": med(upd[:, 15]), "updater_done": med(upd[:, 6]),
           "wait_after_own_last_header": med(upd[:, 12] - np.nanmax(wk[:, :, 5], axis=1)),
           "headers_to_command": med(upd[:, 15] - upd[:, 12]), "command_to_done": med(upd[:, 6] - upd[:, 15])}
    every = [None] * world
    dist.all_gather_object(every, out)
    if rank == 0:
        print(json.dumps({"workload": w.name, "K_per_gpu": w.K, "T": w.T, "world": world, "l2_flushed": not a.no_flush,
                          "ranks": every}))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--math", default="strict")
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--K", type=int, default=0)
    ap.add_argument("--T", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between launches")
    ap.add_argument("--dump", default="", help="save the per-block stamps of every repetition as .npy [reps, nblocks, 8]")
    ap.add_argument("--sharded", action="store_true",
                    help="run under torchrun: the sample-sharded step (fused NVLink exchange), one summary per rank; the "
                         "traced launch is the last of 8 back-to-back steps, so the ranks are in their steady-state lock step")
    a = ap.parse_args()
    if a.sharded:
        return sharded(a)

    import torch
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state

    w = syn.WORKLOADS[a.workload]
    if a.K or a.T:
        w = dataclasses.replace(w, K=a.K or w.K, T=a.T or w.T)
    dev = torch.device("cuda", 0)
    dem = syn.crater_dem(w.grid_size, w.half_width).to(dev)
    cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
    start, goal = syn.workload_start_goal(w)
    core = Core(w.K, w.T, math=a.math,
                variant={"auto": capi.VARIANT_AUTO, "mono": capi.VARIANT_MONO, "pipe": capi.VARIANT_PIPE}[a.variant])
    core.set_terrain(dem, w.half_width, cm)
    st = make_state(start[0], start[1], goal_x=goal[0], goal_y=goal[1])
    nb = C.c_int32()
    capi.check(core.L.mppi_set_trace(core.h, None, C.byref(nb)), "mppi_set_trace")
    trace = torch.zeros((nb.value, NS), dtype=torch.int64, device=dev)
    capi.check(core.L.mppi_set_trace(core.h, trace.data_ptr(), None), "mppi_set_trace")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(5):
        core.step(st, capi.PROJ_3D, None, 42, i)
    torch.cuda.synchronize()
    rows, evt, raw_rows, clk_rows = [], [], [], []
    for i in range(a.reps):
        if not a.no_flush:
            flush.fill_(i & 255)
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        core.step(st, capi.PROJ_3D, None, 42, 100 + i)
        e1.record()
        torch.cuda.synchronize()
        evt.append(e0.elapsed_time(e1) * 1e3)
        t = trace.cpu().numpy().astype(np.float64)
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        rel[t == 0] = np.nan
        rel[:, 7] = t[:, 7]
        rows.append(rel)
        raw_rows.append(rel)
        clk_rows.append(t.copy())
    if a.dump:
        np.save(a.dump, np.stack(raw_rows))
    rel_all = np.stack(rows)                              # [reps, grid.x, 32]
    pipe = bool(np.isfinite(rel_all[:, -1, 12]).all() and not np.isfinite(rel_all[:, -1, 3]).any())
    # pipelined kernel (LL protocol): the last row is the UPDATER block, the others are worker blocks
    rel = rel_all[:, :-1] if pipe else rel_all
    out = {"workload": w.name, "K": w.K, "T": w.T, "math": a.math, "variant": a.variant,
           "nblocks": int(rel.shape[1]), "updater_block": pipe,
           "event_us": {"median": float(np.median(evt)), "min": float(np.min(evt))},
           "sms_used": int(len(np.unique(trace[:, 7].cpu().numpy()))), "phases_us": {}}
    import warnings
    warnings.simplefilter("ignore")
    worker_phases = {0: "entry", 1: "setup_done", 2: "rollout_start", 3: "rollout_end", 4: "roles_joined", 8: "cost_ready",
                     5: "header_published" if pipe else "partial_published", 10: "block_softmax_done",
                     11: "rows_published" if pipe else "ticket"}
    for j, name in worker_phases.items():
        x = rel[:, :, j]
        out["phases_us"][name] = {"first": float(np.nanmedian(np.nanmin(x, axis=1))),
                                  "median": float(np.nanmedian(x)),
                                  "last": float(np.nanmedian(np.nanmax(x, axis=1)))}
    if pipe:
        upd = rel_all[:, -1]
        for j, name in {0: "entry", 1: "poll_start", 2: "headers_complete", 12: "g_min", 13: "g_fold", 14: "g_nominal",
                        15: "g_cmd", 6: "update_done"}.items():
            out["phases_us"]["updater_" + name] = float(np.nanmedian(upd[:, j]))
        out["tail_us"] = {
            "last_header_published_to_headers_complete": float(np.nanmedian(upd[:, 2] - np.nanmax(rel[:, :, 5], axis=1))),
            "headers_complete_to_command": float(np.nanmedian(upd[:, 15] - upd[:, 2])),
            "command_to_update_done": float(np.nanmedian(upd[:, 6] - upd[:, 15])),
            "last_rollout_end_to_update_done": float(np.nanmedian(upd[:, 6] - np.nanmax(rel[:, :, 3], axis=1)))}
        rawu = np.stack(clk_rows)[:, -1]                      # SM-clock stamps of the updater's real pass
        names = {17: "own_headers", 18: "block_min", 19: "scales_compacted", 20: "S_sums", 21: "fold_loaded",
                 22: "fold_barrier", 23: "nominal", 24: "command_stats", 26: "recurrence(warp1)", 27: "recurrence_barrier",
                 28: "outputs"}
        out["updater_cycles_since_poll_start"] = {nm: float(np.median(rawu[:, sl] - rawu[:, 16])) for sl, nm in names.items()}
    else:
        for j, name in {12: "g_min", 13: "g_fold", 14: "g_nominal", 15: "g_cmd", 6: "update_done"}.items():
            out["phases_us"][name] = float(np.nanmedian(np.nanmax(rel[:, :, j], axis=1)))
        # SM-clock stamps of the last block's update (cycles since its ticket was requested)
        raw = np.stack(clk_rows)                              # [reps, nblocks, 32]
        cyc = {}
        names = {23: "ticket_known", 16: "enter_update", 17: "min_done", 18: "scales_compacted", 19: "fold_done",
                 20: "nominal_done", 21: "recurrence_done", 22: "outputs_done"}
        for slot, name in names.items():
            vals = []
            for r in range(raw.shape[0]):
                b = int(np.argmax(raw[r, :, 22]))             # the block that ran the update
                if raw[r, b, 22] > 0 and raw[r, b, slot] > 0:
                    vals.append(raw[r, b, slot] - raw[r, b, 24])
            if vals:
                cyc[name] = float(np.median(vals))
        out["last_block_update_cycles"] = cyc
    x = rel[:, :, 25] - rel[:, :, 2]
    out["phases_us"]["tile_landed_after_rollout_start"] = {"min": float(np.nanmin(x)), "median": float(np.nanmedian(x)),
                                                           "max": float(np.nanmax(x))}
    # is slowness tied to the block (= its samples), to the SM, or random?  correlation of per-block duration across reps
    durb = rel[:, :, 3] - rel[:, :, 25]
    z = durb - np.nanmean(durb, axis=1, keepdims=True)
    cc = np.corrcoef(z[::2].mean(0), z[1::2].mean(0))[0, 1]
    out["block_duration_corr_even_vs_odd_reps"] = float(cc)
    slow = np.argsort(-np.nanmean(durb, axis=0))[:6]
    out["slowest_blocks"] = [{"block": int(b), "smid": int(rel[0, b, 7]), "mean_us": float(np.nanmean(durb[:, b]))} for b in slow]
    fast = np.argsort(np.nanmean(durb, axis=0))[:6]
    out["fastest_blocks"] = [{"block": int(b), "smid": int(rel[0, b, 7]), "mean_us": float(np.nanmean(durb[:, b]))} for b in fast]
    dur = rel[:, :, 3] - rel[:, :, 2]
    out["rollout_us_per_block"] = {"min": float(np.nanmin(dur)), "median": float(np.nanmedian(dur)),
                                   "max": float(np.nanmax(dur)), "per_step_ns_median": float(np.nanmedian(dur) * 1e3 / w.T)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
