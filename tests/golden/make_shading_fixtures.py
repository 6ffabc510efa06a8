"""Generates tests/golden/shading_<config>.npz: high-spp renders of BASELINE.json's configs by the CPU oracle (the f64
restatement of Raytracer.GetColor), stored as sums over 8x8-pixel tiles. The GPU shading-parity tests compare their converged
f32 images against these (north_star: per-tile means within a Monte Carlo bound, whole-image RMSE within 1 %).

  python tests/golden/make_shading_fixtures.py [c1 c3 c4 ...] [--threads N]

Runs the oracle only (no GPU). Minutes to tens of minutes of CPU per config; the outputs are small and committed.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TILE = 8
# name: (scene factory args, width, height, recursion, oracle spp, seed); spp is >= 16x what the GPU test's config-spp pass uses
CONFIGS = {
    # BASELINE C1: Scenes/bounce.txt at 512x512, 64 spp, max 8 bounces -> oracle at 64 x 64 spp
    "c1": dict(file="cornell_bounce.scene", width=512, height=512, recursion=8, spp=4096, seed=21),
    # BASELINE C2 (die.txt, DOF, recursion 3) at reduced size
    "c2": dict(file="die.scene", width=256, height=128, recursion=3, spp=2048, seed=22),
    # BASELINE C3: 1 M-triangle soup, recursion 4, at reduced resolution (the scene is the full-size one)
    "c3": dict(synth="soup", n=1_000_000, sseed=0xC3, jitter=0.01, width=256, height=256, recursion=4, spp=256, seed=23),
    # BASELINE C4: 100 k spheres mirror / glass / diffuse (Fresnel, TIR), recursion 8, at reduced resolution
    "c4": dict(synth="spheres", n=100_000, sseed=0xC4, jitter=0.0, width=256, height=128, recursion=8, spp=4096, seed=24),
    # the same with the oracle's self-hit rule switched to the f32 mode's (oracle/rtc_oracle.h, orc_set_selfhit_mode): on this
    # scene the reference's image depends on f64 rounding noise (paths trapped inside the tiny spheres by re-hits 1e-11 from
    # their origin), which no f32 arithmetic can reproduce; this fixture isolates that one documented deviation
    "c4f": dict(synth="spheres", n=100_000, sseed=0xC4, jitter=0.0, width=256, height=128, recursion=8, spp=4096, seed=24, selfhit=1),
}


def make_scene(cfg):
    from raytracercore_b200 import Scene
    if "synth" in cfg:
        sc = Scene.synthetic(cfg["synth"], cfg["n"], cfg["sseed"], cfg["jitter"])
    else:
        sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", cfg["file"]))
    sc.override(width=cfg["width"], height=cfg["height"], recursion=cfg["recursion"])
    return sc


def tile_sums(a, tile=TILE):
    h, w = a.shape[:2]
    assert h % tile == 0 and w % tile == 0
    return a.reshape((h // tile, tile, w // tile, tile) + a.shape[2:]).sum(axis=(1, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=sorted(CONFIGS))
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    import oracle as O
    for name in a.configs:
        cfg = CONFIGS[name]
        sc = make_scene(cfg)
        O.set_selfhit_mode(cfg.get("selfhit", 0))
        ora = O.OracleScene(sc, seed=cfg["seed"])
        W, H = cfg["width"], cfg["height"]
        acc = (np.zeros((H, W, 3)), np.zeros((H, W), np.uint32), np.zeros((H, W), np.uint32))
        # two independent halves (even / odd sample blocks) give the fixture's own noise level
        half = [(np.zeros((H, W, 3)), np.zeros((H, W), np.uint32), np.zeros((H, W), np.uint32)) for _ in range(2)]
        t = time.time()
        rays = 0
        step = max(1, cfg["spp"] // 16)
        for k, s0 in enumerate(range(0, cfg["spp"], step)):
            _, _, _, r = ora.render(s0, step, threads=a.threads, accum=half[k & 1])
            rays += r
            print("  %s: %d/%d spp, %.0f s" % (name, s0 + step, cfg["spp"], time.time() - t), flush=True)
        rgb = half[0][0] + half[1][0]
        s = half[0][1] + half[1][1]
        m = half[0][2] + half[1][2]
        out = os.path.join(HERE, "shading_%s.npz" % name)
        np.savez_compressed(out, tile=TILE, width=W, height=H, recursion=cfg["recursion"], spp=cfg["spp"], seed=cfg["seed"],
                            rgb=tile_sums(rgb), samples=tile_sums(s.astype(np.int64)), misses=tile_sums(m.astype(np.int64)),
                            rgb_half0=tile_sums(half[0][0]), samples_half0=tile_sums(half[0][1].astype(np.int64)),
                            rays=rays, selfhit=cfg.get("selfhit", 0))
        print("%s: %d rays in %.0f s -> %s (%d bytes)" % (name, rays, time.time() - t, out, os.path.getsize(out)), flush=True)
        ora.close()
        O.set_selfhit_mode(0)


if __name__ == "__main__":
    main()
