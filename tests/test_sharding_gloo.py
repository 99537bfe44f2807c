"""CPU, world_size = 2, gloo: the sample-sharded protocol.  Each rank produces the softmax partial
of its contiguous sample range from the ORACLE's costs (the product's partial comes from the CUDA kernel and is
tested on the GPU box), exchanges it with the product's `exchange_partials`, and folds the gathered partials.
Checks: every rank ends with the identical update, equal to the unsharded one; Philox noise is keyed by the
global sample id so the per-sample costs do not depend on the number of ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, T, lam, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c as oc
    from mppi_b200.sharding import exchange_partials, shard_range
    from util import default_state, terrain
    dem, cm, hw = terrain("small")
    st = default_state(x=-5.0, y=-4.0, goal_x=8.0, goal_y=9.0)
    k0, kl = shard_range(K, world, rank)
    e1, e2 = oc.philox_normals(7, 3, kl, T, k0=k0)           # global sample id keys the stream
    nom = np.full(T, 0.4, np.float32)
    r = oc.mppi_step(oc.make_params(K=kl, T=T, lam=lam), dem, hw, cm, st, nom, nom, e1, e2,
                     dump=["cost", "u1", "u2"])
    cost, u1, u2 = r.dump["cost"], r.dump["u1"], r.dump["u2"]
    m = cost.min()
    w = np.exp(-(cost.astype(np.float64) - m) / lam)
    part = np.zeros(4 + 2 * T, np.float32)
    part[0], part[1], part[3] = m, w.sum(), (w ** 2).sum()
    part[2] = np.array([k0 + int(np.argmin(cost))], np.int32).view(np.float32)[0]
    part[4:4 + T] = (w[:, None] * u1).sum(0)
    part[4 + T:] = (w[:, None] * u2).sum(0)
    allp = exchange_partials(torch.from_numpy(part)).numpy()
    packed = np.concatenate([allp[:, :3], allp[:, 4:]], axis=1)
    n1, n2, M, arg, S = oc.combine_partials(packed, T, lam)
    q.put((rank, n1, n2, M, arg, cost, k0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lam", [0.3, 500.0])
def test_two_rank_sample_sharding_matches_unsharded(oracle, lam):
    K, T, world = 512, 30, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, T, lam, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded reference
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import default_state, terrain
    dem, cm, hw = terrain("small")
    st = default_state(x=-5.0, y=-4.0, goal_x=8.0, goal_y=9.0)
    e1, e2 = oracle.philox_normals(7, 3, K, T)
    nom = np.full(T, 0.4, np.float32)
    ref = oracle.mppi_step(oracle.make_params(K=K, T=T, lam=lam), dem, hw, cm, st, nom, nom, e1, e2, dump=["cost"])
    # costs do not depend on the sharding
    assert np.array_equal(np.concatenate([r[5] for r in res]), ref.dump["cost"])
    # every rank holds the identical result (bitwise), equal to the unsharded update within rounding
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert res[0][4] == res[1][4] == ref.argmin and res[0][3] == ref.min_cost
    np.testing.assert_allclose(res[0][1], ref.nominal1_f64, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(res[0][2], ref.nominal2_f64, rtol=2e-6, atol=1e-7)
