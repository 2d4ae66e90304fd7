"""Multi-GPU sharding of generate_image (reference src/main.rs:94-104).

Every (pixel, sample) path is independent; the only coupling is the per-pixel SUM over samples and
the non-linear tonemap after it.  So the image is partitioned by SAMPLE RANGES: rank r renders a
disjoint range of global sample indices of every pixel into its own f64 accumulation buffer, the
buffers are sum-reduced to rank 0 (one NCCL reduce over NVLink, W*H*3 doubles), and only rank 0
tonemaps.  The RNG stream of a path is keyed by (pixel, GLOBAL sample index), so the union of the
ranks' work is exactly the single-GPU job; only the order of the final per-pixel additions differs.
"""


def sample_range(rank, world, total_samples, first_sample=0):
    """Contiguous block partition of [first_sample, first_sample + total_samples): sizes differ by
    at most one, earlier ranks take the larger blocks.  Returns (begin, count)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total_samples, world)
    count = base + (1 if rank < extra else 0)
    begin = first_sample + rank * base + min(rank, extra)
    return begin, count


def step_sample_range(rank, world, step, spp_per_step):
    """Weak-scaling bench step k: every rank renders spp_per_step samples, globally disjoint."""
    return (step * world + rank) * spp_per_step, spp_per_step


def row_band(rank, world, width, height, align=4):
    """Tile sharding (the other partition north_star names): rank r renders a band of whole rows at
    the full sample count.  Returns the region (x0, y0, w, h) for Context.render_accum, or None
    when there are more ranks than row groups.  Bands are `align` rows granular (a warp of the
    megakernel covers 8x4 pixels).  The bands are disjoint and every rank's buffer is zero outside
    its band, so the same sum-reduce assembles the image — bit for bit the single-GPU one, since no
    pixel receives two contributions.  Load balance is what it costs: a band of sky finishes long
    before a band of terrain, which is why samples are the default partition."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    groups = (height + align - 1) // align
    g0, gn = sample_range(rank, world, groups)
    if gn == 0:
        return None
    y0 = g0 * align
    return 0, y0, width, min(height, (g0 + gn) * align) - y0


def reduce_accum(accum, dst=0, group=None):
    """Sum-reduce the accumulation buffers onto `dst` (in place on dst).  `accum` is a torch tensor
    viewing the library's device buffer (NCCL) or any CPU tensor (gloo in the tests).  Must run
    BEFORE the tonemap: ACES is non-linear and the reference tonemaps the mean (src/main.rs:104)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum
