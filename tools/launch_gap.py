#!/usr/bin/env python
"""Where do the ~6 us between the CUDA events around a C2 step and the kernel's own first / last stamp go?
Stream order: [L2 flush] ts0 | step | ts1, with ts = a one-thread kernel storing %globaltimer (the clock of the
kernel-internal stamps).  Prints medians of: ts0 -> first block's entry, updater's last stamp -> ts1, and the same
bracket around an EMPTY launch (ts0 | ts | ts1) for the cost of a launch boundary itself.  Diagnostic (GPU needed)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state
    flush_on = "--no-flush" not in sys.argv
    dev = torch.device("cuda", 0)
    w = syn.WORKLOADS["C2"]
    dem = syn.crater_dem(w.grid_size, w.half_width).to(dev)
    cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
    start, goal = syn.workload_start_goal(w)
    core = Core(w.K, w.T)
    core.set_terrain(dem, w.half_width, cm)
    st = make_state(start[0], start[1], goal_x=goal[0], goal_y=goal[1])
    L = core.L
    nb = C.c_int32()
    capi.check(L.mppi_set_trace(core.h, None, C.byref(nb)), "trace")
    trace = torch.zeros((nb.value, 32), dtype=torch.int64, device=dev)
    capi.check(L.mppi_set_trace(core.h, trace.data_ptr(), None), "trace")
    ts = torch.zeros(4, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    for i in range(5):
        core.step(st, capi.PROJ_3D, None, 42, i)
    torch.cuda.synchronize()
    rows = []
    for i in range(30):
        if flush_on:
            flush.fill_(i)
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.mppi_test_timestamp(ts.data_ptr(), s)
        e0.record()
        core.step(st, capi.PROJ_3D, None, 42, 100 + i)
        e1.record()
        L.mppi_test_timestamp(ts.data_ptr() + 8, s)
        # empty bracket
        L.mppi_test_timestamp(ts.data_ptr() + 16, s)
        L.mppi_test_timestamp(ts.data_ptr() + 24, s)
        torch.cuda.synchronize()
        t = trace.cpu().numpy().astype(np.float64)
        tt = ts.cpu().numpy().astype(np.float64)
        entry = t[:, 0][t[:, 0] > 0].min()
        done = t[-1, 6]
        rows.append([(entry - tt[0]) / 1e3, (tt[1] - done) / 1e3, (done - entry) / 1e3, (tt[1] - tt[0]) / 1e3,
                     e0.elapsed_time(e1) * 1e3, (tt[3] - tt[2]) / 1e3])
    r = np.median(np.array(rows), axis=0)
    print(json.dumps({"l2_flushed": flush_on, "ts0_to_first_block_entry_us": r[0], "updater_done_to_ts1_us": r[1],
                      "first_entry_to_updater_done_us": r[2], "ts0_to_ts1_us": r[3], "event_us": r[4],
                      "empty_ts_to_ts_us": r[5]}))


if __name__ == "__main__":
    main()
