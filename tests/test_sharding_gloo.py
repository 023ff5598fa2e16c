"""CPU, world_size 2, gloo: the N > 1 host logic of bench.py — shard plan, sum buffers, reduce to rank 0, finalize —
with the oracle standing in for the renderer (the CUDA back end needs a GPU; its own shard/sum test is
tests/test_gpu_render.py::test_sample_sharded_sum_equals_single_gpu)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_plan_covers_every_seed_once():
    from pupiloptixlab_b200 import shard
    for world in (1, 2, 4, 8):
        seen = []
        for step in range(3):
            for r in range(world):
                p = shard.plan(r, world, step, 5)
                assert p.total_spp == 5 * world and p.seed_stride == world
                seen += p.seeds()
        assert sorted(seen) == list(range(3 * 5 * world))
    with pytest.raises(ValueError):
        shard.plan(2, 2, 0, 1)


def test_strong_plan_splits_a_fixed_job():
    """pb2_shard_plan, strong: the step's sample count is fixed and split; ragged counts and more ranks than samples included"""
    from pupiloptixlab_b200 import shard
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 5, 8, 64, 1024):
            seen = []
            for step in range(2):
                parts = [shard.plan(r, world, step, spp, strong=True) for r in range(world)]
                assert all(p.total_spp == spp for p in parts) and sum(p.spp for p in parts) == spp
                assert max(p.spp for p in parts) - min(p.spp for p in parts) <= 1
                for p in parts:
                    seen += p.seeds()
            assert sorted(seen) == list(range(2 * spp))


def _worker(rank, world, port, out_path):
    sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
    import torch
    import torch.distributed as dist
    import orc
    from pupiloptixlab_b200 import scenes, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = scenes.cornell_box(24, 24, 5)
    sc = orc.OracleScene(orc.port(), desc)
    total = None
    for step in range(2):
        p = shard.plan(rank, world, step, 3)
        sums = np.zeros((24 * 24, 4), np.float32)
        for s in p.seeds():
            sums += sc.render(1, first_seed=s, accumulate=False, threads=1)["frame"]
        t = torch.from_numpy(sums)
        shard.reduce_sums(t, dist)
        if rank == 0:
            total = t.numpy().copy() if total is None else total + t.numpy()
    if rank == 0:
        np.save(out_path, total[:, :3] / (2 * 3 * world))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    import orc
    from pupiloptixlab_b200 import scenes
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "mean.npy"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    sharded = np.load(out)
    single = orc.OracleScene(orc.port(), scenes.cornell_box(24, 24, 5)).render(12)["accum"][:, :3]  # seeds 0..11, running mean
    assert np.allclose(sharded, single, rtol=2e-5, atol=1e-6)
