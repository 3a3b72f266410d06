"""Known answers of the reference's `remove_small_regions` (sam2_train/utils/amg.py:246-282, OpenCV 8-connected
components) on seeded masks -> tests/golden/amg_regions.npz.  Build container only; tests read the committed file."""
import os
import sys

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
sys.path.insert(0, "/root/reference")
import make_golden  # noqa: E402  (hydra / omegaconf are absent here: `sam2_train/__init__.py` imports them)

make_golden._stub_hydra()
from sam2_train.utils.amg import remove_small_regions  # noqa: E402


def masks():
    g = np.random.default_rng(17)
    yy, xx = np.mgrid[0:96, 0:128]
    out = []
    for _ in range(4):
        m = np.zeros((96, 128), bool)
        for _ in range(5):                                  # a few discs ...
            cy, cx, r = g.uniform(0, 96), g.uniform(0, 128), g.uniform(4, 22)
            m |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        m ^= g.random((96, 128)) < 0.01                     # ... with salt-and-pepper islands and holes
        out.append(m)
    out.append(np.zeros((96, 128), bool))
    out.append(np.ones((96, 128), bool))
    tiny = np.zeros((96, 128), bool); tiny[3:5, 3:5] = True; tiny[40, 40] = True     # only small islands
    out.append(tiny)
    return out


def main():
    res = {}
    for i, m in enumerate(masks()):
        res[f"mask_{i}"] = m
        for mode in ("holes", "islands"):
            for thr in (5, 40):
                o, changed = remove_small_regions(m, thr, mode)
                res[f"out_{i}_{mode}_{thr}"] = np.asarray(o, bool)
                res[f"changed_{i}_{mode}_{thr}"] = np.array(changed)
    res["n"] = np.array(len(masks()))
    np.savez_compressed(os.path.join(OUT, "amg_regions.npz"), **res)
    print("regions", len(res))


if __name__ == "__main__":
    main()
