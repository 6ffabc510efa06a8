"""Golden data from the reference's own renders: Screenshots/bounce-with-lens.png (Scenes/bounce.txt, 1200x1200) and
Screenshots/die.png (Scenes/die.txt, 1280x960) are the only outputs of the real RaytracerCore that exist (it has no tests
and cannot run here). They are RGBA: FullRaytracer.GetBitmap / SampleSet.GetOutput (SampleSet.cs:61-113) write the gamma-2.2
colour of the hit samples and alpha = 1 - misses / (samples + misses), so alpha is the exact silhouette of the scene through
the scene file's camera and rgb the converged radiance at exposure 1.

This script reduces each screenshot to 20x20-pixel block means (alpha, and rgb premultiplied by alpha, both in [0,1]) and
stores them as a small fixture; tests/test_screenshots.py renders the same scene and compares block by block. Run in the
authoring container (needs /root/reference and Pillow):  python tests/golden/make_screenshot_fixture.py
"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SHOTS = "/root/reference/Screenshots"
BLOCK = 20


def blocks(path):
    a = np.asarray(Image.open(path).convert("RGBA"), dtype=np.float64) / 255.0
    h, w, _ = a.shape
    assert h % BLOCK == 0 and w % BLOCK == 0
    alpha = a[..., 3]
    pre = a[..., :3] * alpha[..., None]
    bh, bw = h // BLOCK, w // BLOCK
    al = alpha.reshape(bh, BLOCK, bw, BLOCK).mean(axis=(1, 3))
    pr = pre.reshape(bh, BLOCK, bw, BLOCK, 3).mean(axis=(1, 3))
    return al.astype(np.float32), pr.astype(np.float32)


if __name__ == "__main__":
    out = {}
    for key, name in (("bounce", "bounce-with-lens.png"), ("die", "die.png")):
        al, pr = blocks(os.path.join(SHOTS, name))
        out[key + "_alpha"] = al
        out[key + "_rgb"] = pr
        print(key, al.shape, "mean alpha %.4f" % al.mean(), "mean premultiplied rgb", pr.reshape(-1, 3).mean(0))
    np.savez_compressed(os.path.join(HERE, "screenshots.npz"), block=np.int32(BLOCK), **out)
    print("wrote", os.path.join(HERE, "screenshots.npz"), os.path.getsize(os.path.join(HERE, "screenshots.npz")), "bytes")
