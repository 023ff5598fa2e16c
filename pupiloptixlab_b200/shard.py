"""Sample-index sharding of a progressive render across the GPUs of one box.

Samples are independent given (pixel, random_seed), so rank r of N renders the seeds
    first_seed(r) + k * N,   k = 0 .. spp_per_rank - 1,      first_seed(r) = base + r
into a private buffer of plain SUMS (`accumulate = 2` / PTPass::SetSumMode: the reference's running mean,
main.cu:190-196, is order dependent), then one reduce(SUM) to rank 0 and a division by the total sample count.
On GPUs the product does all of this itself (PTPass::SetShard, pb2_comm_reduce_frames: NCCL loaded by libpb2.so); this module
is the same plan for host-side tests, where gloo stands in for NCCL and the oracle for the renderer.
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


def plan(rank: int, world: int, step: int, spp: int, strong: bool = False) -> ShardPlan:
    """Seeds of `rank` for progressive step `step`, from the product's own pb2_shard_plan (include/pb2.h; a host-only call).
    weak: every rank renders `spp` frames per step; strong: the step's `spp` frames are split over the ranks.  Over all ranks
    and steps the seeds 0, 1, 2, ... are each rendered exactly once."""
    import ctypes as C
    from . import pb2
    out = [C.c_uint32() for _ in range(4)]
    if pb2.lib().pb2_shard_plan(rank, world, step, spp, int(strong), *[C.byref(o) for o in out]) != 0:
        raise ValueError("bad shard")
    return ShardPlan(first_seed=out[0].value, seed_stride=out[1].value, spp=out[2].value, total_spp=out[3].value)


def reduce_sums(sum_tensor, dist=None, dst: int = 0):
    """In-place reduce(SUM) of the per-rank sum buffers to `dst` (no-op for a single process)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sum_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return sum_tensor
