"""Multi-GPU partitioning of the MPPI step (one process per GPU, torch.distributed for the plumbing).

Two modes, both new work (the reference is single-GPU):

* sample sharding (BASELINE config 3): rank g owns the contiguous global samples [k_begin, k_begin + K_local).
  The Philox counter is the GLOBAL sample id, so the noise -- and therefore every cost -- is independent of the
  number of ranks.  The only exchange is the softmax partial {M, S, argmin, S2, A1[T], A2[T]} (4 + 2T floats per
  rank, 816 B at T = 100); every rank folds the partials in rank order, so all ranks hold the identical updated
  nominal without a broadcast.  Two transports:
    "p2p"  (default on GPUs) the exchange is fused into the step's single launch: every worker block stores its softmax
           partial as flag-in-data lines {value, sequence} into every rank over NVLink (buffers mapped with CUDA IPC) as
           soon as it has it; every rank's updater block polls the world x nblocks slots in its own memory and folds
           them in global block order -- no collective call, no second kernel, no fence, bitwise the result of the
           unsharded launch over the same blocks (many-block launches: the rank folds first, rank partials are exchanged);
    "nccl" one all_gather_into_tensor + the combine kernel (also the path the CPU/gloo tests cover).
* rover sharding (BASELINE config 4): rovers are independent controllers; rank g owns a contiguous block of
  rovers and nothing is exchanged.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` units: returns (begin, count).  The first (total % world) ranks
    get one extra unit; empty shards are possible only when total < world."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def exchange_partials(mine: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather of one softmax partial per rank -> [world, stride] in rank order.  Works on CUDA tensors
    (NCCL over NVLink) and on CPU tensors (gloo; used by the CPU tests of the protocol)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, mine.numel()), dtype=mine.dtype, device=mine.device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine.reshape(-1), group=group)
    else:
        chunks = [out[r] for r in range(world)]
        dist.all_gather(chunks, mine.reshape(-1), group=group)
    return out


class SampleShardedStepper:
    """One logical controller with K_total samples spread over the ranks of `group`."""

    def __init__(self, core, K_total: int, group=None, transport: str = "p2p"):
        self.core, self.group = core, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.k_begin, k_local = shard_range(K_total, self.world, self.rank)
        if k_local != core.K:
            raise ValueError(f"core was created for K={core.K} but this rank's shard has {k_local} samples")
        if transport not in ("p2p", "nccl"):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        self.transport = transport
        n = core.partial_floats()
        self.mine = torch.zeros(n, dtype=torch.float32, device=core.device)
        self.all = torch.zeros((self.world, n), dtype=torch.float32, device=core.device)
        if transport == "p2p":
            self._connect_peers()

    def _connect_peers(self):
        """Exchange the CUDA IPC handles of the per-rank exchange buffers and map every peer's buffer."""
        mine = torch.frombuffer(bytearray(self.core.comm_export(self.world)), dtype=torch.uint8).to(self.core.device)
        if self.world > 1:
            every = torch.empty(64 * self.world, dtype=torch.uint8, device=self.core.device)
            dist.all_gather_into_tensor(every, mine, group=self.group)
        else:
            every = mine
        self.core.comm_connect(self.rank, self.world, bytes(every.cpu().numpy().tobytes()))
        if self.world > 1:
            dist.barrier(group=self.group)

    def step_host(self, state, proj, seed: int, offset: int, stream=None):
        """The sharded step with the command (v*, w*) returned on the host (p2p transport: zero-copy store + poll)."""
        if self.transport == "p2p":
            return self.core.step_sharded_host(state, self.k_begin, proj, seed, offset, stream)
        self.step(state, proj, seed, offset, None, stream)
        v = self.core.stats[0, 6:8].cpu()
        return float(v[0]), float(v[1])

    def step(self, state, proj, seed: int, offset: int, noise=None, stream=None):
        if self.transport == "p2p":
            self.core.step_sharded(state, self.k_begin, proj, noise, seed, offset, stream)
            return
        self.core.step_partial(state, self.mine, self.k_begin, proj, noise, seed, offset, stream)
        if self.world > 1:
            exchange_partials(self.mine, self.group, self.all)
            parts = self.all
        else:
            parts = self.mine
        self.core.combine_partials(state, parts, self.world, stream)
