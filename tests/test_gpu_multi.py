"""Multi-GPU behind the C ABI (include/rt_b200.h "multi-GPU"): the in-library NCCL communicator,
rt_reduce_accum, and the single-process device group.  The partition is by global SAMPLE ranges
(reference src/main.rs:94-103: paths are independent, the only coupling is the per-pixel sum), so
N devices must reproduce the one-device accumulation up to the order of the final additions.

Tests that need two devices skip on a one-GPU box; they run under `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest

import rtb200
import rtb200.ctypes_types as T
from tests import helpers

pytestmark = pytest.mark.gpu


def _scene(tmp_path, w=96, h=64, spp=10):
    path, _ = helpers.heightfield(tmp_path, 30, 40)
    return rtb200.Scene.load_gltf(path, w, h, spp)


def test_group_of_one_is_the_plain_context(tmp_path):
    """rt_group_* on one device needs no NCCL and gives rt_render's image bit for bit (also in chunks)."""
    sc = _scene(tmp_path)
    ctx = rtb200.Context(0).upload(sc)
    want = ctx.render(seed=7)
    ctx.close()
    g = rtb200.Group([0]).upload(sc)
    assert g.size == 1
    assert g.render(seed=7).tobytes() == want.tobytes()
    rgb, done = g.render_progressive(4, seed=7)
    assert done == 10 and rgb.tobytes() == want.tobytes()
    g.close()


def test_world_of_one_communicator(tmp_path):
    """rt_comm_init with world = 1 goes through ncclCommInitRank and ncclReduce inside the library
    (the in-place reduce over one rank is the identity): exercises the NCCL path on a one-GPU box."""
    sc = _scene(tmp_path)
    ctx = rtb200.Context(0).upload(sc)
    ctx.render_accum(0, 4, 11, clear=True)
    want = ctx.accum()
    ctx.reduce_accum(0)                                 # no communicator yet: a world of one, no-op
    assert ctx.accum().tobytes() == want.tobytes()
    ctx.comm_init(rtb200.comm_unique_id(), 0, 1)
    info = ctx.comm_info()
    assert info["world"] == 1 and info["rank"] == 0 and info["nccl_version"] >= 20000
    ctx.reduce_accum(0)
    assert ctx.accum().tobytes() == want.tobytes()
    with pytest.raises(rtb200.RtError):
        ctx.reduce_accum(1)                             # root outside the world
    with pytest.raises(rtb200.RtError):
        ctx.comm_init(rtb200.comm_unique_id(), 0, 1)    # already has one
    ctx.close()


def _close(a, b, rtol=1e-12):
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err[a == b] = 0
    return float(err.max())


@pytest.mark.skipif(rtb200.device_count() < 2, reason="needs two CUDA devices")
def test_group_of_two_equals_one_device(tmp_path):
    """Single-process group: samples split by rt_shard_sample_range, ncclCommInitAll + ncclReduce
    inside the library.  Each device's share is bit for bit the one-device render of that sample
    range, so the reduced buffer equals the sum of the two shares exactly, and the one-device
    accumulation up to the order of the last addition per pixel."""
    sc = _scene(tmp_path, spp=9)
    one = rtb200.Context(0).upload(sc)
    one.render_accum(0, 9, 5, clear=True)
    whole = one.accum()
    parts = []
    for r in range(2):
        b, c = rtb200.shard_sample_range(r, 2, 9)
        one.render_accum(b, c, 5, clear=True)
        parts.append(one.accum())
    want_rgb = one.render(seed=5)
    one.close()
    g = rtb200.Group([0, 1]).upload(sc)
    g.render_accum(0, 9, 5, clear=True)
    g.reduce_accum()
    got = g.ctx(0).accum()
    assert got.tobytes() == (parts[0] + parts[1]).tobytes()
    assert _close(got, whole) <= 1e-12
    rgb = g.render(seed=5)
    assert np.abs(rgb - want_rgb).max() <= 1e-12
    rgb2, done = g.render_progressive(4, seed=5)
    assert done == 9 and np.abs(rgb2 - want_rgb).max() <= 1e-12
    g.close()


def _rank_worker(rank, world, port, gltf, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)       # plumbing only: carries the 128-byte id
    sc = rtb200.Scene.load_gltf(gltf, 96, 64, 9)
    ctx = rtb200.Context(rank).upload(sc)
    box = [rtb200.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
    b, c = rtb200.shard_sample_range(rank, world, 9)
    ctx.render_accum(b, c, 5, clear=True)
    np.save(os.path.join(out_dir, f"part{rank}.npy"), ctx.accum())
    ctx.reduce_accum(0)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), ctx.accum())
        np.save(os.path.join(out_dir, "rgb.npy"), ctx.postprocess(9))
    else:
        ctx.sync()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(rtb200.device_count() < 2, reason="needs two CUDA devices")
def test_two_ranks_through_the_c_abi(tmp_path):
    """One process per GPU (the bench's layout): rt_comm_unique_id -> broadcast -> rt_comm_init ->
    rt_render_accum on the rank's sample range -> rt_reduce_accum -> rt_postprocess on rank 0."""
    import torch.multiprocessing as mp
    gltf, _ = helpers.heightfield(tmp_path, 30, 40)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rank_worker, args=(2, port, gltf, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(str(tmp_path), f"part{r}.npy")) for r in range(2)]
    total = np.load(os.path.join(str(tmp_path), "sum.npy"))
    assert total.tobytes() == (parts[0] + parts[1]).tobytes()
    one = rtb200.Context(0).upload(rtb200.Scene.load_gltf(gltf, 96, 64, 9))
    one.render_accum(0, 9, 5, clear=True)
    assert _close(total, one.accum()) <= 1e-12
    assert np.abs(np.load(os.path.join(str(tmp_path), "rgb.npy")) - one.render(seed=5)).max() <= 1e-12
    one.close()
