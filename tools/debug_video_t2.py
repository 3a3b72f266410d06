"""Debug helper: per-frame error of the 2-object hiera_t video case against the golden fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import medsam2_b200
from test_gpu_e2e import _build, _run_video, G

z = np.load(f"{G}/video_hiera_t_512_2obj.npz")
with medsam2_b200.compute(torch.bfloat16):
    m = _build("sam2_hiera_t", video=True, image_size=512)
    st, outs = _run_video(m, 512, 6, 2, (0, 3), ((3, 1),), 77)
od = st["output_dict"]
for f in range(6):
    o = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)
    ref = z[f"pred_masks_{f}"]
    got = o["pred_masks"].float().cpu().numpy()
    d = np.abs(got - ref)
    filled = (np.abs(got - 0.1) < 1e-6) | (np.abs(ref - 0.1) < 1e-6)
    for ob in range(2):
        dd = d[ob][~filled[ob]]
        print(f"frame {f} obj {ob}: max {dd.max():.4f} mean {dd.mean():.5f} ref range [{ref[ob].min():.2f},{ref[ob].max():.2f}] "
              f"sign agree {((got[ob] > 0) == (ref[ob] > 0)).mean():.4f} ptr err {np.abs(o['obj_ptr'][ob].float().cpu().numpy() - z[f'obj_ptr_{f}'][ob]).max():.4f}")
