"""Sample-index sharding of a progressive render across the GPUs of one box.

Samples are independent given (pixel, random_seed), so rank r of N renders the seeds
    first_seed(r) + k * N,   k = 0 .. spp_per_rank - 1,      first_seed(r) = base + r
into a private buffer of plain SUMS (`accumulate = 2` / PTPass::SetSumMode: the reference's running mean,
main.cu:190-196, is order dependent), then one reduce(SUM) to rank 0 and a division by the total sample count.
No collective runs inside the data path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    first_seed: int
    seed_stride: int
    spp: int          # frames this rank renders in the step
    total_spp: int    # frames all ranks render in the step

    def seeds(self):
        return [self.first_seed + k * self.seed_stride for k in range(self.spp)]


def plan(rank: int, world: int, step: int, spp_per_rank: int) -> ShardPlan:
    """Seeds of `rank` for progressive step `step` (weak scaling: every rank renders spp_per_rank frames per step).
    Over all ranks and steps the seeds 0, 1, 2, ... are each rendered exactly once."""
    if not (0 <= rank < world) or spp_per_rank < 1:
        raise ValueError("bad shard")
    return ShardPlan(first_seed=step * spp_per_rank * world + rank, seed_stride=world, spp=spp_per_rank, total_spp=spp_per_rank * world)


def reduce_sums(sum_tensor, dist=None, dst: int = 0):
    """In-place reduce(SUM) of the per-rank sum buffers to `dst` (no-op for a single process)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sum_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return sum_tensor
