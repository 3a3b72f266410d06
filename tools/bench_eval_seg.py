"""Device time of the one-pass segmentation-metric kernel (ms2_seg_counts, 5 thresholds) on a 96-slice 1024^2 volume
(805 MB read, > L2) against the HBM roofline, and wall time of eval_seg_frames against the reference-style torch
statement of func_3d/utils.py:184-202 run on the same GPU (per-threshold binarise + D2H + numpy/torch reductions)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import numpy as np
import torch
from medsam2_b200 import ops
from medsam2_b200.utils.eval import eval_seg_frames

thr = (0.1, 0.3, 0.5, 0.7, 0.9)
n, hw = 96, 1024
g = torch.Generator().manual_seed(0)
gt = (torch.rand(n, 1, hw, hw, generator=g) > 0.7).float().cuda()
pred = gt * 4 - 1.5 + torch.randn(n, 1, hw, hw, device="cuda")
p2, g2 = pred.view(n, -1), gt.view(n, -1)
for _ in range(3):
    ops.seg_counts(p2, g2, thr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ops.seg_counts(p2, g2, thr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = 2 * 4 * n * hw * hw
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(f"ms2_seg_counts 96 x 1024^2, 5 thresholds: {ms:.3f} ms per launch (memset included), {nbytes / ms / 1e6:.0f} GB/s algorithmic "
      f"(8 B/pixel); MEASURED_PEAKS: {peaks}")

for _ in range(3):
    ops.bce_logits_sum(p2, g2, 2.0)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ops.bce_logits_sum(p2, g2, 2.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"ms2_bce_logits_sum 96 x 1024^2 (validation loss, pos_weight 2): {ms:.3f} ms per launch, {nbytes / ms / 1e6:.0f} GB/s algorithmic (8 B/pixel)")
if "--kernel-only" in sys.argv:
    sys.exit(0)
t0 = time.perf_counter(); ours = eval_seg_frames(pred, gt, thr); t_ours = time.perf_counter() - t0


def ref_style(pred, mask):   # the reference's sequence of tensor ops, c == 1 branch, written out plainly
    eiou = edice = 0
    for th in thr:
        gt_v = (mask > th).float(); vp = (pred > th).float(); vpc = vp.cpu()
        a = vpc[:, 0].numpy().astype("int32"); b = gt_v[:, 0].cpu().numpy().astype("int32")
        inter = (a & b).sum((1, 2)); union = (a | b).sum((1, 2))
        eiou += ((inter + 1e-6) / (union + 1e-6)).mean()
        x, y = vp[0, 0].reshape(-1), gt_v[0, 0].reshape(-1)
        edice += ((2 * torch.dot(x, y) + 0.0001) / (x.sum() + y.sum() + 0.0001)).item()
    return eiou / len(thr), edice / len(thr)


t0 = time.perf_counter(); theirs = [ref_style(pred[i:i + 1], gt[i:i + 1]) for i in range(n)]; t_ref = time.perf_counter() - t0
same = all(abs(a[0] - b[0]) < 1e-12 and abs(a[1] - b[1]) < 1e-6 for a, b in zip(ours, theirs))
print(f"eval of a 96-slice volume: eval_seg_frames {t_ours * 1e3:.1f} ms wall vs reference-style per-slice loop on the same GPU "
      f"{t_ref * 1e3:.1f} ms wall ({t_ref / t_ours:.0f}x); results agree: {same}; mean IoU {np.mean([o[0] for o in ours]):.4f}")
