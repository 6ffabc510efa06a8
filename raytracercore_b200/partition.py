"""Multi-GPU partition of a frame: every (pixel, sample) is independent given the replicated scene
and the Philox stream is keyed by (pixel, sample), so the ranks of a job split each frame's samples of every pixel and one
sum-reduction of the SampleSet planes per frame (rtc_reduce_accum) gives exactly the single-GPU frame (colour sums to f64
rounding, counters bit-exact). Two ways to split, both used by bench.py:

  weak    every rank adds `spp` samples per frame: a frame has world * spp samples (per-GPU work fixed as GPUs are added);
  strong  a frame has `total_spp` samples whatever the number of ranks; rank r takes a contiguous share, the first
          total_spp % world ranks one sample more (total work fixed).

`run_frame` is the per-frame sequence every rank executes; it is the whole multi-GPU data path.
"""


def sample_range(frame, rank, world, spp):
    """Weak split: first sample index and count rendered by `rank` in frame number `frame`."""
    return (frame * world + rank) * spp, spp


def frame_samples(frame, world, spp):
    """Weak split: the samples one frame adds to every pixel across all ranks, [first, first + count)."""
    return frame * world * spp, world * spp


def strong_sample_range(frame, rank, world, total_spp):
    """Strong split: first sample index and count of `rank` when every frame has total_spp samples in all."""
    base, extra = divmod(total_spp, world)
    count = base + (1 if rank < extra else 0)
    first = frame * total_spp + rank * base + min(rank, extra)
    return first, count


def run_frame(ctx, frame, rank, world, spp, root=0, strong=False, clear=False):
    """One frame on one rank: optional clear, this rank's sample range, then the single per-frame collective.
    spp is per rank (weak) or per frame (strong). Returns (first_sample, count) rendered by this rank."""
    first, count = strong_sample_range(frame, rank, world, spp) if strong else sample_range(frame, rank, world, spp)
    if clear:
        ctx.clear_accum()
    if count:
        ctx.render(first, count)
    if world > 1:
        ctx.reduce_accum(root)
    return first, count
