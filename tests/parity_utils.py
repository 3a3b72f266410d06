"""Parity harness for the video path at the benchmarked geometry (hiera_s, 1024², bbox every 2 slices).

Checker = the oracle (`oracle/sam2_oracle.py`, fp32 torch statement of the reference, pinned against the real
reference by tests/golden) run on the same GPU.  Two regimes, as SURVEY App. A.6 asks:

* teacher-forced: every tracked frame of the product is fed the ORACLE's memories (conditioning + recent
  `maskmem_features`, object pointers), so each frame's error is that of ONE pass through memory attention,
  SAM heads, hole filling and the memory encoder — no feedback of sign flips through the binarised memory mask;
* free-running: the product's own `propagate_in_video`, errors accumulate over the volume.

Hole filling writes the constant +0.1 into small background holes; a logit within 1e-3 of zero can flip a
pixel in or out of a hole, so pixels that either side filled are excluded from the max-abs figures and
counted separately (`filled_mismatch`)."""
import torch

from oracle.config import get_config
from oracle.sam2_oracle import OracleSAM2, OracleVideoPredictor
from oracle.weights import make_state_dict
from synth_data import btcv_volume


def _err(a, b, exclude_fill=True):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    if exclude_fill:
        filled = ((a - 0.1).abs() < 1e-6) | ((b - 0.1).abs() < 1e-6)
        mism = int((((a - 0.1).abs() < 1e-6) != ((b - 0.1).abs() < 1e-6)).sum().item())
        d = d.masked_fill(filled, 0.0)
    else:
        mism = 0
    return float(d.max().item()), mism


def oracle_run(cfg_name, size, T, every, seed=1234, autocast=False, fill_hole_area=8):
    """Free-running oracle on cuda: -> (state, {f: video-res logits}, cond temp outputs, volume, boxes)."""
    cfg = get_config(cfg_name, image_size=size)
    vp = OracleVideoPredictor(OracleSAM2(cfg, make_state_dict(cfg), device="cuda"), fill_hole_area=fill_hole_area)
    vol, boxes = btcv_volume(T, size, seed, 1)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with torch.no_grad(), ctx:
        st = vp.init_state(vol.cuda(), size, size)
        for f in range(0, T, every):
            vp.add_new_bbox(st, f, 1, boxes[f][0], clear_old_points=False)
        vid = {f: mk.float().clone() for f, _, mk in vp.propagate_in_video(st, start_frame_idx=0)}
    return vp, st, vid, vol, boxes


def build_product(cfg_name, size, T):
    import medsam2_b200
    m = medsam2_b200.build_sam2_video_predictor(
        cfg_name, device="cuda", hydra_overrides_extra=[f"++model.image_size={size}", f"++model.feature_cache_size={T}",
                                                       "++model.feature_encode_batch=8"])
    m.load_state_dict(make_state_dict(get_config(cfg_name)), strict=True)
    return m


def product_free_run(m, vol, boxes, size, T, every):
    st = m.val_init_state(imgs_tensor=vol.cuda(), video_height=size, video_width=size)
    for f in range(0, T, every):
        m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
    vid = {f: mk.float().clone() for f, _, mk in m.propagate_in_video(st, start_frame_idx=0)}
    return st, vid


def _frame_out(od, f):
    return od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)


def compare_free_running(ost, ovid, pst, pvid, T):
    """-> dict of worst errors over all frames: low-res logits, video-res logits, sign agreement."""
    worst = {"low_res": 0.0, "video_res": 0.0, "sign_agree_min": 1.0, "flipped_abs_logit_max": 0.0, "filled_mismatch": 0,
             "obj_ptr": 0.0}
    for f in range(T):
        o, p = _frame_out(ost["output_dict"], f), _frame_out(pst["output_dict"], f)
        e, mm = _err(p["pred_masks"], o["pred_masks"])
        worst["low_res"] = max(worst["low_res"], e)
        worst["filled_mismatch"] += mm
        # video-res = bilinear x4 of the (hole-filled) low-res logits: exclude the footprint of filled pixels by comparing
        # only where both low-res planes are free of the fill value in the 4x4 neighbourhood; simpler and strict: compare
        # the video-res planes directly when no pixel was filled on either side, else away from 0.1-valued neighbourhoods
        a, b = pvid[f].float(), ovid[f].float()
        fill = (((p["pred_masks"].float() - 0.1).abs() < 1e-6) | ((o["pred_masks"].float() - 0.1).abs() < 1e-6)).float()
        if fill.any():
            infl = torch.nn.functional.max_pool2d(fill, 3, 1, 1)
            infl = torch.nn.functional.interpolate(infl, size=a.shape[-2:], mode="nearest") > 0
            d = (a - b).abs().masked_fill(infl, 0.0)
        else:
            d = (a - b).abs()
        worst["video_res"] = max(worst["video_res"], float(d.max().item()))
        flip = (a > 0) != (b > 0)
        worst["sign_agree_min"] = min(worst["sign_agree_min"], 1.0 - float(flip.float().mean().item()))
        # a pixel may only change sides where the oracle's own logit is within the error bound of zero (random-weight
        # logits crowd around zero, so the FRACTION of such pixels says little; their magnitude is the criterion)
        flip_free = flip & ~infl if fill.any() else flip
        if flip_free.any():
            worst["flipped_abs_logit_max"] = max(worst["flipped_abs_logit_max"], float(b.abs()[flip_free].max().item()))
        worst["obj_ptr"] = max(worst["obj_ptr"], float((p["obj_ptr"].float() - o["obj_ptr"].float()).abs().max().item()))
    return worst


def teacher_forced(m, ovp, ost, vol, boxes, size, T, every):
    """Per-frame errors of the product with the oracle's memories fed in (see module docstring).
    -> dict: worst low-res / video-res logit error over tracked frames, cond frames, maskmem features, pointers."""
    import medsam2_b200  # noqa: F401
    from medsam2_b200.utils.misc import fill_holes_in_mask_scores  # noqa: F401
    ood = ost["output_dict"]
    res = {"cond_low_res": 0.0, "cond_obj_ptr": 0.0, "cond_maskmem": 0.0, "tracked_low_res": 0.0, "tracked_video_res": 0.0,
           "tracked_obj_ptr": 0.0, "tracked_maskmem": 0.0, "filled_mismatch": 0, "per_frame_low_res": {}}
    with torch.inference_mode():
        st = m.val_init_state(imgs_tensor=vol.cuda(), video_height=size, video_width=size)
        prompts = list(range(0, T, every))
        for f in prompts:
            m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
            p = st["temp_output_dict_per_obj"][0]["cond_frame_outputs"][f]
            o = ood["cond_frame_outputs"][f]
            e, mm = _err(p["pred_masks"], o["pred_masks"])
            res["cond_low_res"] = max(res["cond_low_res"], e)
            res["filled_mismatch"] += mm
            res["per_frame_low_res"][f] = e
            res["cond_obj_ptr"] = max(res["cond_obj_ptr"], float((p["obj_ptr"].float() - o["obj_ptr"].float()).abs().max().item()))
            # memory encoder of the prompted frame on the ORACLE's consolidated mask (sam2_video_predictor.py:746-862)
            high = torch.nn.functional.interpolate(o["pred_masks"].float(), size=(size, size), mode="bilinear", align_corners=False)
            mf, _ = m._run_memory_encoder(st, f, 1, high.contiguous(), True)
            res["cond_maskmem"] = max(res["cond_maskmem"], float((mf.float() - o["maskmem_features"].float()).abs().max().item()))
        forced = {"cond_frame_outputs": dict(ood["cond_frame_outputs"]), "non_cond_frame_outputs": {}}
        for f in range(T):
            if f in forced["cond_frame_outputs"]:
                continue
            out, pred = m._run_single_frame_inference(
                inference_state=st, output_dict=forced, frame_idx=f, batch_size=1, is_init_cond_frame=False,
                point_inputs=None, mask_inputs=None, reverse=False, run_mem_encoder=True)
            o = ood["non_cond_frame_outputs"][f]
            e, mm = _err(out["pred_masks"], o["pred_masks"])
            res["tracked_low_res"] = max(res["tracked_low_res"], e)
            res["per_frame_low_res"][f] = e
            res["filled_mismatch"] += mm
            if mm == 0:
                _, vid = m._get_orig_video_res_output(st, pred)
                ovid = torch.nn.functional.interpolate(o["pred_masks"].float(), size=(size, size), mode="bilinear", align_corners=False)
                fill = ((o["pred_masks"].float() - 0.1).abs() < 1e-6).float()
                d = (vid.float() - ovid).abs()
                if fill.any():
                    infl = torch.nn.functional.interpolate(torch.nn.functional.max_pool2d(fill, 3, 1, 1), size=(size, size)) > 0
                    d = d.masked_fill(infl, 0.0)
                res["tracked_video_res"] = max(res["tracked_video_res"], float(d.max().item()))
            res["tracked_obj_ptr"] = max(res["tracked_obj_ptr"],
                                         float((out["obj_ptr"].float() - o["obj_ptr"].float()).abs().max().item()))
            res["tracked_maskmem"] = max(res["tracked_maskmem"],
                                         float((out["maskmem_features"].float() - o["maskmem_features"].float()).abs().max().item()))
            forced["non_cond_frame_outputs"][f] = o           # the NEXT frames see the oracle's memory of this one
    return res
