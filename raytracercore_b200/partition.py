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


def check_reduced_frame(job, spp):
    """bench.py's in-run check at N > 1 (collective: every rank calls it). One weak-split frame is rendered by all ranks and
    reduced to rank 0; rank 0 then renders the same world * spp samples of every pixel alone. The reduced planes must equal
    the single-GPU ones: `samples` and `misses` bit for bit (and samples + misses == world * spp everywhere), the f64 colour
    sums to 1e-12 relative (the summation order differs between the two). Returns a small report on rank 0, None elsewhere."""
    import numpy as np
    env, ctx = job.env, job.ctx
    rank, world = env["rank"], env["world"]
    ctx.clear_accum()
    run_frame(ctx, 0, rank, world, spp)
    got = ctx.read_accum() if rank == 0 else None
    ctx.sync()
    env["barrier"]()
    report = None
    if rank == 0:
        ctx.clear_accum()
        ctx.render(*frame_samples(0, world, spp))
        w_rgb, w_s, w_m = ctx.read_accum()
        rgb, s, m = got
        exact = bool(np.array_equal(s, w_s) and np.array_equal(m, w_m))
        total_ok = bool(np.all(s.astype(np.int64) + m == world * spp))
        scale = np.maximum(np.abs(w_rgb), 1e-300)
        rel = float(np.max(np.abs(rgb - w_rgb) / scale)) if rgb.size else 0.0
        close = bool(np.allclose(rgb, w_rgb, rtol=1e-12, atol=1e-12))
        report = {"result": "green" if (exact and total_ok and close) else "red", "counters_bit_exact": exact,
                  "samples_plus_misses_equal_n_spp": total_ok, "max_rel_diff_rgb_sum": rel, "samples_per_pixel": world * spp,
                  "what": "reduced planes of one frame over N ranks vs rank 0's own render of the same N x spp samples"}
    ctx.clear_accum()
    env["barrier"]()
    return report
