"""World-size-2 CPU (gloo) test of the multi-GPU frame partition (raytracercore_b200/partition.py): each rank renders its
sample range with the oracle standing in for the device, one sum-reduce of the SampleSet planes per frame, and the
result must equal the single-rank frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp, frames, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    from raytracercore_b200 import Scene
    from raytracercore_b200.partition import sample_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", "cornell_bounce.scene"))
    sc.override(width=20, height=12, recursion=4)
    ora = O.OracleScene(sc, seed=9)
    total = None
    for f in range(frames):
        first, n = sample_range(f, rank, world, spp)
        rgb, s, m, _ = ora.render(first, n, threads=1)
        planes = [torch.from_numpy(rgb), torch.from_numpy(s.astype(np.int64)), torch.from_numpy(m.astype(np.int64))]
        for p in planes:  # the one collective per frame
            dist.reduce(p, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            total = planes if total is None else [a + b for a, b in zip(total, planes)]
    if rank == 0:
        np.savez(out_path, rgb=total[0].numpy(), s=total[1].numpy(), m=total[2].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_partition_equals_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    from raytracercore_b200 import Scene
    from raytracercore_b200.partition import frame_samples, sample_range
    world, spp, frames = 2, 2, 2
    # the ranges of all ranks tile the frame's sample interval exactly
    for f in range(3):
        first, count = frame_samples(f, world, spp)
        got = sorted(s for r in range(world) for s in range(*(lambda a, n: (a, a + n))(*sample_range(f, r, world, spp))))
        assert got == list(range(first, first + count))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "frame.npz")
    mp.spawn(_worker, args=(world, port, spp, frames, out), nprocs=world, join=True)
    got = np.load(out)
    sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", "cornell_bounce.scene"))
    sc.override(width=20, height=12, recursion=4)
    rgb, s, m, _ = O.OracleScene(sc, seed=9).render(0, world * spp * frames, threads=1)
    assert np.array_equal(got["s"], s) and np.array_equal(got["m"], m)
    assert np.allclose(got["rgb"], rgb, rtol=1e-12, atol=1e-12)


def test_strong_split_tiles_every_frame_exactly():
    from raytracercore_b200.partition import strong_sample_range
    for world in (1, 2, 3, 4, 8):
        for total in (1, 4, 7, 8, 16, 4096):
            for f in range(3):
                got = []
                for r in range(world):
                    first, n = strong_sample_range(f, r, world, total)
                    got += list(range(first, first + n))
                assert got == list(range(f * total, (f + 1) * total)), (world, total, f)
