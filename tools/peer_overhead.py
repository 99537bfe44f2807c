#!/usr/bin/env python
"""What does peer memory cost a launch?  Under torchrun (2+ GPUs): CUDA-event time of (a) the plain step before any
peer mapping exists, (b) the same plain step after the exchange buffers of the other ranks have been mapped (CUDA IPC),
(c) the sharded step that stores into them, (d) an empty kernel-sized torch op, before / after.  Diagnostic only."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def p50(fn, n=300):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        ev[i][0].record(); fn(i); ev[i][1].record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]) * 1e3)


def main():
    from mppi_b200 import capi, synthetic as syn
    from mppi_b200.core import Core, make_state
    from mppi_b200.sharding import SampleShardedStepper
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = syn.WORKLOADS["C2"]
    dem = syn.crater_dem(w.grid_size, w.half_width).to(dev)
    cm = torch.from_numpy(syn.rock_costmap(w.costmap_size, w.half_width)).to(dev)
    start, goal = syn.workload_start_goal(w)
    st = make_state(start[0], start[1], goal_x=goal[0], goal_y=goal[1])
    plain = Core(w.K, w.T, device=local)
    plain.set_terrain(dem, w.half_width, cm)
    x = torch.zeros(1024, device=dev)
    out = {"rank": rank}
    p50(lambda i: plain.step(st, capi.PROJ_3D, None, 42, i), 50)
    out["plain_before_mapping_us"] = p50(lambda i: plain.step(st, capi.PROJ_3D, None, 42, i))
    out["tiny_op_before_us"] = p50(lambda i: x.add_(1.0))
    core = Core(w.K, w.T, device=local)
    core.set_terrain(dem, w.half_width, cm)
    stepper = SampleShardedStepper(core, w.K * world, transport="p2p")
    out["plain_after_mapping_us"] = p50(lambda i: plain.step(st, capi.PROJ_3D, None, 42, i))
    out["tiny_op_after_us"] = p50(lambda i: x.add_(1.0))
    dist.barrier()
    p50(lambda i: stepper.step(st, capi.PROJ_3D, 42, i), 50)
    dist.barrier()
    out["sharded_us"] = p50(lambda i: stepper.step(st, capi.PROJ_3D, 42, 1000 + i))
    every = [None] * world
    dist.all_gather_object(every, out)
    if rank == 0:
        print(json.dumps(every))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
