"""Generate tests/golden/eval_seg_cases.npz by running the REAL reference metric
(/root/reference/func_3d/utils.py `eval_seg`) on seeded inputs.  Build container only
(`python tests/golden/make_golden_eval.py`); also the validation loss `BCEWithLogitsLoss(pos_weight=2)` of
func_3d/function.py:35-36 per example and over the batch.  Tests read the committed .npz, never /root/reference."""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.argv = [sys.argv[0]]                      # func_3d/utils.py parses the command line at import (cfg.parse_args)
sys.path.insert(0, "/root/reference")
import func_3d.utils as ref_utils  # noqa: E402

THRESHOLDS = (0.1, 0.3, 0.5, 0.7, 0.9)        # func_3d/function.py:216


def blob_mask(g, b, c, h, w):
    """soft-edged random ellipses in [0,1] (ground truth masks are 0/1 in BTCV; soft values exercise every threshold)"""
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    out = torch.zeros(b, c, h, w)
    for i in range(b):
        for j in range(c):
            cy, cx = torch.rand(2, generator=g) * torch.tensor([h, w])
            ry, rx = (0.1 + 0.3 * torch.rand(2, generator=g)) * torch.tensor([h, w])
            d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
            out[i, j] = torch.clamp(1.5 - d, 0, 1)
    return out


def main():
    g = torch.Generator().manual_seed(1234)
    cases = {}
    specs = [("c1_b1", 1, 1, 64, 64, THRESHOLDS), ("c1_b3", 3, 1, 48, 80, THRESHOLDS), ("c2_b2", 2, 2, 64, 64, THRESHOLDS),
             ("c1_odd", 2, 1, 37, 53, THRESHOLDS), ("c1_empty", 2, 1, 32, 32, THRESHOLDS), ("c1_full", 1, 1, 32, 32, (0.5,)),
             ("c3_b2_one_thr", 2, 3, 40, 40, (0.5,)), ("c1_hard01", 2, 1, 128, 128, THRESHOLDS)]
    for name, b, c, h, w, thr in specs:
        gt = blob_mask(g, b, c, h, w)
        pred = gt * 4 - 1.5 + torch.randn(b, c, h, w, generator=g)          # logits (func_3d/function.py:283 keeps them raw)
        if name == "c1_empty":
            gt.zero_()
            pred = pred - 100.0                                             # nothing above any threshold: smooth terms only
        if name == "c1_full":
            gt.fill_(1.0)
            pred = pred.abs() + 1.0
        if name == "c1_hard01":
            gt = (gt > 0.5).float()
        res = ref_utils.eval_seg(pred, gt, thr)
        cases[name + "/pred"] = pred.numpy()
        cases[name + "/gt"] = gt.numpy()
        cases[name + "/thr"] = np.array(thr, dtype=np.float64)
        cases[name + "/res"] = np.array([float(r) for r in res], dtype=np.float64)
        # the validation loss exactly as func_3d/function.py:35-36,299 builds and calls it (pos_weight = 2), CPU fp32
        crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.ones([1]) * 2)
        cases[name + "/bce"] = np.array([float(crit(pred[i:i + 1], gt[i:i + 1])) for i in range(b)] + [float(crit(pred, gt))],
                                        dtype=np.float64)
        print(name, [float(r) for r in res])
    np.savez_compressed(os.path.join(OUT, "eval_seg_cases.npz"), **cases)


if __name__ == "__main__":
    main()
