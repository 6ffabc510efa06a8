"""Multi-GPU partition of a frame: every (pixel, sample) is independent given the replicated scene and
the Philox stream is keyed by (pixel, sample), so rank r of R renders the sample range
[first + r*spp, first + (r+1)*spp) of every pixel and one sum-reduction of the SampleSet planes per frame gives
exactly the single-GPU frame of R*spp samples (sums to f64 rounding, counters bit-exact)."""


def sample_range(frame, rank, world, spp):
    """First sample index and count rendered by `rank` in frame number `frame`."""
    return (frame * world + rank) * spp, spp


def frame_samples(frame, world, spp):
    """The samples one frame adds to every pixel across all ranks: [first, first + count)."""
    return frame * world * spp, world * spp
