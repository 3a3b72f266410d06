"""Segmentation metrics of the validation loop — the step after the hot path.

Mirror of the reference's `func_3d/utils.py` (`eval_seg` :139-202, `iou` :204-214, `dice_coeff` :215-240; call site
`func_3d/function.py:276-305`), same names, arguments and return values.  The reference binarises prediction and ground
truth once per threshold, copies both masks to the host and reduces them with numpy / one `torch.dot` per example
(for the default 5 thresholds: 10 full-resolution passes, 10 D2H copies and 5 `.item()` syncs per object and slice).
Here ONE kernel (`ms2_seg_counts`) reads both planes once and returns three exact integers per plane and threshold;
IoU and Dice follow on the host from those integers with the reference's arithmetic (float64 IoU, fp32 Dice), so the
results are bit-identical as long as a plane has < 2^24 pixels (beyond that the REFERENCE's fp32 sums round).

No CPU fallback: tensors must live on the GPU (`ops._chk` raises otherwise).
"""
import numpy as np
import torch

from .. import ops

_SMOOTH = 1e-6                 # func_3d/utils.py:206
_EPS = np.float32(0.0001)      # func_3d/utils.py:231 (enters fp32 tensor arithmetic)


def _planes(x):
    if x.dtype != torch.float32:
        x = x.float()
    return x if x.is_contiguous() else x.contiguous()


def seg_counts(pred, true_mask_p, threshold):
    """pred, true_mask_p [b,c,h,w] on the GPU -> numpy int64 [b,c,T,3] = (inter, |pred>th|, |gt>th|); one launch per
    8 thresholds and ONE device->host copy (3*T integers per plane)."""
    b, c = pred.shape[:2]
    thr = [float(t) for t in threshold]
    p, g = _planes(pred).reshape(b * c, -1), _planes(true_mask_p).reshape(b * c, -1)
    parts = [ops.seg_counts(p, g, thr[i:i + 8]) for i in range(0, len(thr), 8)]
    counts = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    return counts.cpu().numpy().astype(np.int64).reshape(b, c, len(thr), 3)


def _iou_from_counts(k):
    """k int64 [b,3] for one class and threshold -> `iou` (func_3d/utils.py:204-214): float64 mean over the batch."""
    inter = k[:, 0]
    union = k[:, 1] + k[:, 2] - k[:, 0]
    return ((inter + _SMOOTH) / (union + _SMOOTH)).mean()


def _dice_from_counts(k):
    """k int64 [b,3] -> `dice_coeff` (func_3d/utils.py:215-240): fp32 per-example ratio, fp32 running sum, / b."""
    s = np.float32(0.0)
    for inter, np_, ng in k:
        union = np.float32(np.float32(np.float32(np_) + np.float32(ng)) + _EPS)
        t = np.float32(np.float32(np.float32(2.0) * np.float32(inter) + _EPS) / union)
        s = np.float32(s + t)
    return float(np.float32(s / np.float32(len(k))))


def _reduce(counts):
    """counts int64 [b,c,T,3] -> the tuple `eval_seg` returns."""
    b, c, T, _ = counts.shape
    ious, dices = [0] * c, [0] * c
    for t in range(T):
        for i in range(c):
            ious[i] += _iou_from_counts(counts[:, i, t])
            dices[i] += _dice_from_counts(counts[:, i, t])
    if c == 1:
        return ious[0] / T, dices[0] / T
    if c == 2:
        return ious[0] / T, ious[1] / T, dices[0] / T, dices[1] / T
    return tuple(np.array(ious + dices) / T)


def eval_seg(pred, true_mask_p, threshold):
    """Drop-in for `func_3d.utils.eval_seg(pred, true_mask_p, threshold)`: pred / true_mask_p [b,c,h,w] (logits or
    probabilities, compared with `>` against every threshold).  c == 1 -> (iou, dice); c == 2 -> (iou_d, iou_c,
    dice_d, dice_c); c > 2 -> c ious followed by c dices."""
    if pred.dim() != 4 or pred.shape != true_mask_p.shape:
        raise ValueError(f"eval_seg: expected matching [b,c,h,w] tensors, got {tuple(pred.shape)} and "
                         f"{tuple(true_mask_p.shape)}")
    if len(threshold) == 0:
        raise ZeroDivisionError("eval_seg: empty threshold tuple")      # the reference divides by len(threshold)
    return _reduce(seg_counts(pred, true_mask_p, threshold))


def eval_seg_frames(preds, true_masks, threshold):
    """The metric half of the loop at func_3d/function.py:276-305 for a whole volume: preds / true_masks [n,c,h,w] hold
    one (slice, object) pair per row; returns the list of `eval_seg(pred[i:i+1], mask[i:i+1], threshold)` tuples from
    ONE launch and ONE device->host copy instead of n * (10 passes + 10 copies + 5 syncs)."""
    if preds.dim() != 4 or preds.shape != true_masks.shape:
        raise ValueError("eval_seg_frames: expected matching [n,c,h,w] tensors")
    counts = seg_counts(preds, true_masks, threshold)
    return [_reduce(counts[i:i + 1]) for i in range(counts.shape[0])]


def iou(outputs, labels):
    """`func_3d.utils.iou` for GPU tensors of 0/1 values [b,h,w]."""
    k = seg_counts(outputs.unsqueeze(1), labels.unsqueeze(1), (0.5,))
    return _iou_from_counts(k[:, 0, 0])


def dice_coeff(input, target):
    """`func_3d.utils.dice_coeff` for GPU tensors of 0/1 values [b,...]; returns a 1-element fp32 tensor on the
    inputs' device like the reference."""
    b = input.shape[0]
    k = seg_counts(input.reshape(b, 1, 1, -1), target.reshape(b, 1, 1, -1), (0.5,))
    return torch.tensor([_dice_from_counts(k[:, 0, 0])], dtype=torch.float32, device=input.device)


def bce_with_logits_frames(preds, targets, pos_weight=2.0):
    """Validation loss of func_3d/function.py:35-36,299 (`BCEWithLogitsLoss(pos_weight=2)(pred, mask)`, mean reduction) for
    every row of preds / targets [n, ...] at once: fp32 tensor [n] on the device, one launch, no synchronisation."""
    if preds.shape != targets.shape:
        raise ValueError("bce_with_logits_frames: expected tensors of the same shape")
    n = preds.shape[0]
    p, g = _planes(preds).reshape(n, -1), _planes(targets).reshape(n, -1)
    return (ops.bce_logits_sum(p, g, pos_weight) / max(p.shape[1], 1)).float()


def bce_with_logits(pred, target, pos_weight=2.0):
    """`torch.nn.BCEWithLogitsLoss(pos_weight=pos_weight)(pred, target)`: scalar fp32 tensor on the device."""
    if pred.shape != target.shape:
        raise ValueError("bce_with_logits: expected tensors of the same shape")
    return bce_with_logits_frames(pred.reshape(1, -1), target.reshape(1, -1), pos_weight)[0]


def score_frames_lowres(low, true_masks, threshold, pos_weight=2.0):
    """`eval_seg_frames` + `bce_with_logits_frames` of the video-resolution logits WITHOUT materialising them: low
    [n,1,h,w] low-resolution logits (what the tracker keeps per slice), true_masks [n,1,H,W]; the bilinear up-sampling
    of sam2_video_predictor.py:724-744 happens inside the scoring pass (`ms2_score_lowres`, bit-identical to
    `ms2_resize_bilinear` followed by `ms2_seg_counts`).  -> (list of eval_seg tuples, fp32 losses [n] on the device)."""
    if low.dim() != 4 or true_masks.dim() != 4 or low.shape[:2] != true_masks.shape[:2] or low.shape[1] != 1:
        raise ValueError("score_frames_lowres: expected [n,1,h,w] logits and [n,1,H,W] masks")
    n = low.shape[0]
    thr = [float(t) for t in threshold]
    if len(thr) == 0:
        raise ZeroDivisionError("score_frames_lowres: empty threshold tuple")
    lo, gt = _planes(low).reshape(n, *low.shape[2:]), _planes(true_masks).reshape(n, *true_masks.shape[2:])
    parts, sums = [], None
    for i in range(0, len(thr), 8):
        c, s = ops.score_lowres(lo, gt, thr[i:i + 8], pos_weight if i == 0 else None)
        parts.append(c)
        sums = s if i == 0 else sums
    counts = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    counts = counts.cpu().numpy().astype(np.int64).reshape(n, 1, len(thr), 3)
    losses = (sums / float(gt.shape[1] * gt.shape[2])).float()
    return [_reduce(counts[i:i + 1]) for i in range(n)], losses
