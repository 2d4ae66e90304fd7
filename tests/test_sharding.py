"""Multi-GPU host logic on the CPU: sample-range sharding + sum-reduce over a 2-rank gloo group,
with the oracle standing in for the GPU renderer (the reduce and the partition are what is under
test; the device path is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rtb200
from rtb200 import scenes, sharding


def test_sample_range_partitions_exactly():
    for total in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                b, c = sharding.sample_range(r, world, total, first_sample=5)
                covered += list(range(b, b + c))
            assert covered == list(range(5, 5 + total))
    seen = set()
    for k in range(3):
        for r in range(4):
            b, c = sharding.step_sample_range(r, 4, k, 8)
            s = set(range(b, b + c))
            assert not (s & seen)
            seen |= s
    assert seen == set(range(96))
    with pytest.raises(ValueError):
        sharding.sample_range(2, 2, 10)


def test_library_partition_is_the_python_one():
    """rt_shard_sample_range (what rt_group_* and the CLI use) == sharding.sample_range."""
    for total in (0, 1, 7, 64, 1024, 0xFFFFFFF0):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                assert rtb200.shard_sample_range(r, world, total, 5) == sharding.sample_range(r, world, total, 5)
    with pytest.raises(rtb200.RtError):
        rtb200.shard_sample_range(2, 2, 10)


def test_multi_gpu_entry_points_fail_cleanly_without_a_device():
    """No GPU here: the group constructor reports RT_ERR_NO_DEVICE (2), it never falls back to a CPU path;
    the communicator id comes from NCCL itself (dlopen'ed) or is RT_ERR_COMM (8) when NCCL is absent."""
    if rtb200.device_count() == 0:
        with pytest.raises(rtb200.RtError) as e:
            rtb200.Group()
        assert e.value.code == 2
    try:
        uid = rtb200.comm_unique_id()
        assert len(uid) == rtb200.COMM_ID_BYTES and any(uid)
    except rtb200.RtError as e:
        assert e.code == 8


def _worker(rank, world, port, text, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import loaders, orc
    scene = orc.OracleScene(loaders.parse_text_scene(text))
    begin, count = sharding.sample_range(rank, world, total)
    accum, _ = scene.render_accum(begin, count, 0xC0FFEE, 1)
    t = torch.from_numpy(accum)
    sharding.reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_rank(orc, tmp_path):
    from oracle import loaders
    text = scenes.cornell_text(24, 16, 6, 4, scenes.CORNELL_EXTRA)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = os.path.join(str(tmp_path), "acc.npy")
    mp.spawn(_worker, args=(2, port, text, 6, out), nprocs=2, join=True)
    got = np.load(out)
    want, _ = orc.OracleScene(loaders.parse_text_scene(text)).render_accum(0, 6, 0xC0FFEE, 1)
    assert want.sum() > 0
    # same paths, per-pixel sums formed in a different order: equal to rounding
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    # and the tonemapped image agrees
    np.testing.assert_allclose(orc.postprocess(got, 6), orc.postprocess(want, 6), rtol=1e-12, atol=1e-15)


def test_row_bands_partition_the_image():
    """Tile sharding: the bands are disjoint, 4-row granular and cover every row exactly once."""
    from rtb200 import sharding
    for height in (1, 3, 4, 38, 1080, 2160):
        for world in (1, 2, 3, 8):
            rows = []
            for rank in range(world):
                r = sharding.row_band(rank, world, 100, height)
                if r is None:
                    continue
                x0, y0, w, h = r
                assert (x0, w) == (0, 100) and h > 0 and y0 % 4 == 0
                rows += list(range(y0, y0 + h))
            assert rows == list(range(height)), (height, world)
    with pytest.raises(ValueError):
        sharding.row_band(2, 2, 10, 10)


def _band_worker(rank, world, port, text, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import loaders, orc
    raw = loaders.parse_text_scene(text)
    full, _ = orc.OracleScene(raw).render_accum(0, total, 0xC0FFEE, 1)     # the oracle renders whole images only:
    band = sharding.row_band(rank, world, raw.width, raw.height)            # keep this rank's band, zero elsewhere
    mine = np.zeros_like(full)
    if band is not None:
        _, y0, _, h = band
        mine[y0:y0 + h] = full[y0:y0 + h]                                     # (H, W, 3)
    t = torch.from_numpy(mine)
    sharding.reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_two_rank_row_bands_assemble_the_image_exactly(orc, tmp_path):
    """Tile sharding over a 2-rank gloo group: every pixel receives exactly one non-zero
    contribution, so the sum-reduce assembles the single-rank buffer BIT FOR BIT (unlike sample
    ranges, where two partial sums are added)."""
    from oracle import loaders
    text = scenes.cornell_text(24, 18, 3, 4, scenes.CORNELL_EXTRA)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = os.path.join(str(tmp_path), "bands.npy")
    mp.spawn(_band_worker, args=(2, port, text, 3, out), nprocs=2, join=True)
    want, _ = orc.OracleScene(loaders.parse_text_scene(text)).render_accum(0, 3, 0xC0FFEE, 1)
    got = np.load(out)
    assert got.shape == want.shape and got.tobytes() == want.tobytes() and want.sum() > 0
