"""Automatic mask generator on one 1024^2 image with the reference's default grid (32 x 32 points, 64 per batch, 3 masks
per point): wall time of generate() through the public API (image from host memory, records back on the host), and device
time of the statistics kernel on one batch of candidates (192 planes of 1024^2 = 805 MB) against the HBM roofline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import medsam2_b200
from medsam2_b200 import ops
from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
from oracle.config import get_config
from oracle.weights import make_state_dict
from test_amg import amg_image

x = torch.randn(192, 1024, 1024, device="cuda")
for _ in range(3):
    ops.mask_stats(x, 0.0, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.mask_stats(x, 0.0, 1.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"ms2_mask_stats 192 x 1024^2: {ms:.3f} ms per launch, {x.numel() * 4 / ms / 1e6:.0f} GB/s algorithmic (4 B/pixel)")
sel = torch.arange(0, 192, 4, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.mask_binarize_t(x, sel, 0.0, (1024, 1024), (0, 0))
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    ops.mask_binarize_t(x, sel, 0.0, (1024, 1024), (0, 0))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"ms2_mask_binarize_t 48 of 192 planes of 1024^2: {ms:.3f} ms per launch, {48 * 1024 * 1024 * 5 / ms / 1e6:.0f} GB/s algorithmic (4 B in + 1 B out per pixel)")
del x
m = medsam2_b200.build_sam2("sam2_hiera_s", device="cuda")
m.load_state_dict(make_state_dict(get_config("sam2_hiera_s")), strict=True)
img = np.ascontiguousarray(np.kron(amg_image(256, 256), np.ones((4, 4, 1), np.uint8)))
gen = SAM2AutomaticMaskGenerator(m, points_per_side=32, points_per_batch=64, pred_iou_thresh=0.0, stability_score_thresh=0.5,
                                 stability_score_offset=0.02, box_nms_thresh=0.7, output_mode="uncompressed_rle")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    anns = gen.generate(img)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"generate(): 1024 prompts -> 3072 candidates -> {len(anns)} records in {dt * 1e3:.1f} ms wall ({3072 / dt:.0f} candidates/s), hiera_s bf16, random weights")
