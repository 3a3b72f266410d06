import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, R + "/medical-sam2_b200"); sys.path.insert(0, R + "/tests")
import numpy as np, torch
from oracle.config import get_config
from oracle.weights import make_state_dict
from synth_data import btcv_volume
import medsam2_b200
z = np.load(R + "/tests/golden/video_clicks_hiera_t_512.npz")
size, T = 512, 6
for dt in (torch.float32, torch.bfloat16):
    with medsam2_b200.compute(dt):
        m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_t", device="cuda", hydra_overrides_extra=["++model.image_size=512"])
        m.load_state_dict(make_state_dict(get_config("sam2_hiera_t")), strict=True)
        vol, boxes = btcv_volume(T, size, 55, 1)
        st = m.val_init_state(imgs_tensor=vol.cuda(), video_height=size, video_width=size)
        x0, y0, x1, y1 = boxes[2][0]
        c2 = [(x0 + x1) / 2.0, (y0 + y1) / 2.0]
        orig = m._forward_sam_heads_impl
        def spy(*a, **k):
            r = orig(*a, **k)
            print("   ious", r[2].float().cpu().numpy().round(4), "obj", r[6].float().cpu().numpy().round(3))
            return r
        m._forward_sam_heads_impl = spy
        m.use_cuda_graphs = False
        _, _, vr = m.train_add_new_points(inference_state=st, frame_idx=2, obj_id=1, points=torch.tensor([c2]),
                                          labels=torch.tensor([1], dtype=torch.int32), clear_old_points=False)
        d = np.abs(vr[..., ::4, ::4].float().cpu().numpy() - z["a/click1_video_res_sub"])
        print(dt, "click1 err quantiles 50/90/99/99.9/max", [float(np.quantile(d, q)) for q in (0.5, 0.9, 0.99, 0.999, 1.0)],
              "frac>0.01", float((d > 0.01).mean()))
        lo = st["temp_output_dict_per_obj"][0]["cond_frame_outputs"][2]["pred_masks"].float().cpu().numpy()
        print("   low-res filled px", int((np.abs(lo - 0.1) < 1e-6).sum()), "range", lo.min(), lo.max())
        print("second click")
        _, _, vr = m.train_add_new_points(inference_state=st, frame_idx=2, obj_id=1,
                                          points=torch.tensor([[c2[0] + 150.0, c2[1] + 120.0]]),
                                          labels=torch.tensor([0], dtype=torch.int32), clear_old_points=False)
        lo = st["temp_output_dict_per_obj"][0]["cond_frame_outputs"][2]["pred_masks"].float().cpu().numpy()
        d = np.abs(lo - z["a/pred_masks_2"])
        print(dt, "click2 low-res err quantiles 50/90/99/99.9/max", [float(np.quantile(d, q)) for q in (0.5, 0.9, 0.99, 0.999, 1.0)],
              "frac>0.01", float((d > 0.01).mean()))
