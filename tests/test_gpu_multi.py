"""-m gpu: the multi-GPU path inside the product: PTPass::SetShard -> pb2_shard_plan -> plain sums ->
pb2_comm_reduce_frames (NCCL loaded by libpb2.so).  One-rank cases run on any box; the two-rank cases need two GPUs and skip
otherwise (the driver's scaling run exercises 2 / 4 / 8 ranks through bench.py)."""
import subprocess

import numpy as np
import pytest

from pupiloptixlab_b200 import pb2, pupil, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _system():
    pupil.init(0)
    yield
    pupil.shutdown()


def test_one_rank_shard_equals_the_plain_pass():
    """a communicator of one rank needs no NCCL: sums + finalize give the running mean to fp32 rounding, step after step"""
    desc = scenes.material_grid(96, 54, 6)
    pupil.load_scene(desc)
    pupil.pass_config(frames_per_run=12)
    pupil.run(1)
    want = pupil.buffer("final result")[..., :3].astype(np.float64)
    try:
        pupil.set_shard(0, 1, None, strong=True)
        pupil.pass_config(frames_per_run=12)
        pupil.run(1)
        got = pupil.buffer("final result")  # the download waits for the asynchronous reduction
        assert np.all(got[..., 3] == 1.0)
        assert np.allclose(got[..., :3], want, rtol=2e-5, atol=1e-6)
        assert np.all(pupil.buffer("pt accum buffer")[..., 3] == 12)  # w counts the frames summed
        # progressive: 3 runs of 4 frames continue the same sums (seeds 0..11) and the frame is sum / 12 again
        pupil.pass_config(frames_per_run=4)
        pupil.run(3)
        got3 = pupil.buffer("final result")
        assert pupil.pass_state() == (12, 12)
        assert np.allclose(got3[..., :3], want, rtol=2e-5, atol=1e-6)
        # a restart does not read the old sums (accumulate = 2 with sample_cnt = 0 starts over)
        pupil.pass_config(frames_per_run=12)
        pupil.run(1)
        assert np.array_equal(pupil.buffer("final result"), got)
    finally:
        pupil.set_shard(0, 0, None)


def test_reduce_frames_through_the_c_abi_one_rank():
    L = pb2.lib()
    import ctypes as C
    comm = C.c_void_p()
    pb2.check(L.pb2_comm_create(C.byref(comm), 1, 0, None))
    s = pb2.Scene()
    s.build()
    n = 1000
    sums = np.random.default_rng(1).uniform(0, 50, (n, 4)).astype(np.float32)
    d_sum, d_frame = pb2.DeviceBuffer(sums.nbytes), pb2.DeviceBuffer(sums.nbytes)
    d_sum.upload(sums)
    for mode in (0, 1):
        d_frame.zero()
        pb2.check(L.pb2_comm_reduce_frames(comm, s.h, d_sum.ptr, d_frame.ptr, n, 8, mode, 0))
        pb2.check(L.pb2_comm_synchronize(comm))
        out = d_frame.download(np.float32, (n, 4))
        assert np.array_equal(out[:, :3], sums[:, :3] * np.float32(1.0 / 8)) and np.all(out[:, 3] == 1.0)
    assert L.pb2_comm_reduce_frames(comm, s.h, None, d_frame.ptr, n, 8, 0, 0) != 0      # no sums
    assert L.pb2_comm_reduce_frames(comm, s.h, d_sum.ptr, d_frame.ptr, n, 8, 7, 0) != 0  # no such mode
    assert L.pb2_comm_reduce_frames(comm, s.h, d_sum.ptr, d_frame.ptr, n, 8, 0, 3) != 0  # no such root
    bad = C.c_void_p()
    assert L.pb2_comm_create(C.byref(bad), 2, 0, None) != 0  # two ranks need the id
    assert L.pb2_comm_create(C.byref(bad), 2, 5, None) != 0
    pb2.check(L.pb2_comm_destroy(comm))


def _pfm(path):
    raw = path.read_bytes()
    head, rest = raw.split(b"-1.0\n", 1)
    _, dims = head.split(b"\n", 1)
    w, h = (int(x) for x in dims.split())
    return np.frombuffer(rest, np.float32).reshape(h, w, 3)


@pytest.mark.parametrize("spp", [16, 5])
def test_path_tracer_on_two_gpus_matches_one(tmp_path, spp):
    """path_tracer --gpus 2: two processes, NCCL id on the command line, strong split of --spp (ragged for 5), rank 0 writes the
    image: the same picture as one GPU to fp32 rounding (sums in another order than the running mean)"""
    if pb2.lib().pb2_device_count() < 2:
        pytest.skip("needs two GPUs")
    xml = scenes.to_xml(scenes.material_grid(160, 90, 6), tmp_path / "grid.xml")
    exe = pb2.PKG / "_build" / "path_tracer"
    one, two = tmp_path / "one.pfm", tmp_path / "two.pfm"
    r = subprocess.run([str(exe), "--scene", str(xml), "--spp", str(spp), "--batch", "8", "--out", str(one)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), "--scene", str(xml), "--spp", str(spp), "--batch", "8", "--gpus", "2", "--out", str(two)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "on 2 GPUs" in r.stdout
    a, b = _pfm(one), _pfm(two)
    assert np.isfinite(b).all() and np.allclose(a, b, rtol=3e-5, atol=2e-6)
