"""Run under torchrun (one process per GPU): the sample-sharded MPPI step with the exchange fused into the launch
(NVLink peer memory) against the NCCL all-gather transport and against the unsharded controller.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from mppi_b200 import capi
    from mppi_b200.core import Core, make_state
    from mppi_b200.sharding import SampleShardedStepper, shard_range
    from util import terrain
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dem, cm, hw = terrain("C1")
    dem_t, cm_t = torch.from_numpy(dem).to(dev), torch.from_numpy(cm).to(dev)
    st = make_state(-60.57, -60.23, goal_x=65.8, goal_y=65.4)
    results = {}
    # (K_total, T): the latency regime (flat exchange: every block's partial goes to every rank, bitwise the unsharded
    # result) and the throughput regime (thousands of blocks: two-level exchange of one rank partial)
    for K_total, T, lam in ((4096, 60, 0.3), (4096, 60, 2000.0), (32768 * world, 20, 2000.0)):
        nom = np.full(T, 0.35, np.float32)
        flat = (K_total // world) <= 148 * 32
        # reference: the whole problem on this GPU
        full = Core(K_total, T, device=local, lambda_=lam)
        full.set_terrain(dem_t, hw, cm_t)
        full.set_nominal(nom, nom)
        full.step(st, capi.PROJ_3D, None, 5, 11)
        torch.cuda.synchronize()
        ref = (full.optimal_u1[0].cpu().numpy().copy(), full.read_stats())
        full.close()
        _, k_local = shard_range(K_total, world, rank)
        for transport in ("p2p", "nccl"):
            core = Core(k_local, T, device=local, lambda_=lam)
            core.set_terrain(dem_t, hw, cm_t)
            stepper = SampleShardedStepper(core, K_total, transport=transport)
            for it in range(4):                       # several iterations: both parities, re-armed flags
                core.set_nominal(nom, nom)
                stepper.step(st, capi.PROJ_3D, 5, 11)
                torch.cuda.synchronize()
                got = (core.optimal_u1[0].cpu().numpy().copy(), core.read_stats())
                assert got[1]["argmin"] == ref[1]["argmin"], (transport, lam, it, got[1], ref[1])
                assert got[1]["min_cost"] == ref[1]["min_cost"]
                err = float(np.max(np.abs(got[0] - ref[0]) / np.maximum(np.abs(ref[0]), 1e-3)))
                assert err < 1e-5, (transport, lam, it, err)
                if transport == "p2p" and flat:
                    # every block's partial is folded in global block order, exactly as the unsharded launch does
                    assert np.array_equal(got[0], ref[0]), (lam, it)
            results[transport] = got[0]
            # every rank must hold the identical nominal (bitwise): gather and compare
            mine = torch.from_numpy(got[0]).to(dev)
            every = torch.empty((world, T), device=dev)
            dist.all_gather_into_tensor(every, mine)
            assert bool((every == every[0]).all()), (transport, lam)
            core.close()
        # the NCCL transport folds per rank first, then across ranks: same value up to the summation order
        assert np.allclose(results["p2p"], results["nccl"], rtol=1e-5, atol=1e-7)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
