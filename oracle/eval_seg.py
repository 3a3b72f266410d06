"""numpy restatement of the reference's segmentation metrics (the step AFTER the hot path).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for `medsam2_b200.utils.eval` — never shipped.

Parity status: PINNED.  `tests/golden/make_golden_eval.py` imports the real `func_3d.utils.eval_seg` from
/root/reference, runs it on seeded inputs and commits inputs + answers as `tests/golden/eval_seg_cases.npz`;
`tests/test_oracle_golden.py` checks this file against them.  Reference paths below are relative to /root/reference/.
"""
import numpy as np


def iou_np(outputs, labels):
    """func_3d/utils.py:204-214 `iou`: int arrays [b,h,w] -> mean over b of (|a&b| + 1e-6)/(|a|b| + 1e-6), float64."""
    smooth = 1e-6
    inter = (outputs & labels).sum((1, 2))
    union = (outputs | labels).sum((1, 2))
    return ((inter + smooth) / (union + smooth)).mean()


def dice_coeff_np(inp, target):
    """func_3d/utils.py:215-240 `dice_coeff` + `DiceCoeff.forward`, in fp32 like the reference's tensors:
    per example (2*<a,b> + 1e-4)/(sum a + sum b + 1e-4), accumulated into an fp32 scalar and divided by b."""
    eps = np.float32(0.0001)
    s = np.float32(0.0)
    n = 0
    for a, b in zip(inp, target):
        a = a.astype(np.float32).reshape(-1)
        b = b.astype(np.float32).reshape(-1)
        inter = np.float32(np.dot(a.astype(np.float64), b.astype(np.float64)))   # 0/1 values: an exact integer
        union = np.float32(np.float32(np.float32(a.sum(dtype=np.float64)) + np.float32(b.sum(dtype=np.float64))) + eps)
        t = np.float32(np.float32(np.float32(2.0) * inter + eps) / union)
        s = np.float32(s + t)
        n += 1
    return float(np.float32(s / np.float32(n)))


def eval_seg_np(pred, true_mask_p, threshold):
    """func_3d/utils.py:139-202 `eval_seg`.  pred, true_mask_p: float arrays [b,c,h,w]; threshold: iterable of floats.
    c == 1 -> (iou, dice); c == 2 -> (iou_d, iou_c, dice_d, dice_c); c > 2 -> c ious then c dices.
    (The reference's c > 2 branch re-binds `pred` inside its threshold loop, :175, so it only runs with ONE
    threshold; this restatement follows the evident intent for several.)"""
    pred = np.asarray(pred, dtype=np.float32)
    gt = np.asarray(true_mask_p, dtype=np.float32)
    c = pred.shape[1]
    ious = [0] * c
    dices = [0] * c
    for th in threshold:
        th32 = np.float32(th)             # the comparison runs on fp32 tensors: the scalar is rounded to fp32
        vg = (gt > th32).astype(np.float32)
        vp = (pred > th32).astype(np.float32)
        for i in range(c):
            ious[i] += iou_np(vp[:, i].astype("int32"), vg[:, i].astype("int32"))
            dices[i] += dice_coeff_np(vp[:, i], vg[:, i])
    n = len(threshold)
    if c == 1:
        return ious[0] / n, dices[0] / n
    if c == 2:
        return ious[0] / n, ious[1] / n, dices[0] / n, dices[1] / n
    return tuple(np.array(ious + dices) / n)


def bce_with_logits_np(pred, target, pos_weight=2.0):
    """func_3d/function.py:35-36,299: `torch.nn.BCEWithLogitsLoss(pos_weight=2)(pred, mask)`, mean over all elements;
    torch's element formula (1-y)*x + (1+(pw-1)*y)*(log1p(exp(-|x|)) + max(-x,0)), evaluated here in float64."""
    x = np.asarray(pred, dtype=np.float64)
    y = np.asarray(target, dtype=np.float64)
    el = (1 - y) * x + (1 + (pos_weight - 1) * y) * (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, 0))
    return float(el.mean())
