"""Validation step over a loader of volumes — the caller around the hot path (reference `func_3d/function.py:198-314`
`validation_sam`).

Same pack format, prompting schedule, metric definitions and return value as the reference driver:
    pack = {"image": [T,3,H,W] (or [1,T,3,H,W]), "label": {frame: {obj_id: mask}}, "bbox": {frame: {obj_id: box}}
            or "pt" / "p_label": {frame: {obj_id: tensor}}}
    returns (mean loss over the loader, (mean IoU, mean Dice)).
What differs is how a volume is scored: the reference calls `lossfunc` and `eval_seg` once per (slice, object) — per
call 5 thresholds x (2 binarisations + 2 D2H copies + numpy reductions) + one `.item()`; here all (slice, object) rows
of the volume are stacked and scored by ONE launch and two small D2H copies.  By default (`fused_scoring=True`) the
tracker hands over its LOW-resolution logits and `ms2_score_lowres` up-samples them to the video resolution inside the
scoring pass (the fp32 video-resolution logits — 4 MiB per object and slice at 1024² — are never written or re-read; the
metrics are bit-identical to the two-step path, the loss agrees to fp32 rounding); `fused_scoring=False` scores the
video-resolution logits `propagate_in_video` yields (`ms2_seg_counts` + `ms2_bce_logits_sum`).  With `shard=True` the volumes of the loader are dealt round-robin to the ranks of the default process
group (`parallel.shard_volumes`) and the averages are formed with one all-reduce (`parallel.reduce_validation`).
"""
import torch

from .parallel import reduce_validation, shard_volumes
from .utils.eval import bce_with_logits_frames, eval_seg_frames, score_frames_lowres

THRESHOLD = (0.1, 0.3, 0.5, 0.7, 0.9)      # func_3d/function.py:205


def _prompt_volume(net, state, pack, prompt, prompt_frames, obj_list, frame_hw, device):
    """func_3d/function.py:236-267: a click or a box per prompted slice and object; an all-zero mask prompt where the
    object has no annotation on that slice."""
    for f in prompt_frames:
        for obj in obj_list:
            try:
                if prompt == "click":
                    net.train_add_new_points(inference_state=state, frame_idx=f, obj_id=obj,
                                             points=pack["pt"][f][obj].to(device),
                                             labels=pack["p_label"][f][obj].to(device), clear_old_points=False)
                elif prompt == "bbox":
                    net.train_add_new_bbox(inference_state=state, frame_idx=f, obj_id=obj,
                                           bbox=torch.as_tensor(pack["bbox"][f][obj]).to(device), clear_old_points=False)
                else:
                    raise ValueError(f"unknown prompt type {prompt!r} (expected 'click' or 'bbox')")
            except KeyError:
                net.train_add_new_mask(inference_state=state, frame_idx=f, obj_id=obj,
                                       mask=torch.zeros(frame_hw, device=device))


@torch.no_grad()
def score_volume(video_segments, mask_dict, frame_ids, obj_list, threshold=THRESHOLD, pos_weight=2.0):
    """func_3d/function.py:276-305 for one volume: (mean loss, mean IoU, mean Dice) over all (slice, object) pairs.
    video_segments[frame][obj] are logits [1,H,W] on the GPU; a missing ground-truth mask counts as all background."""
    preds, masks = [], []
    for f in frame_ids:
        for obj in obj_list:
            p = video_segments[f][obj]
            try:
                m = mask_dict[f][obj].to(dtype=torch.float32, device=p.device).reshape(p.shape)
            except KeyError:
                m = torch.zeros_like(p, dtype=torch.float32)
            preds.append(p.float())
            masks.append(m)
    preds, masks = torch.stack(preds), torch.stack(masks)                    # [n,1,H,W]
    losses = bce_with_logits_frames(preds, masks, pos_weight)                # one launch, stays on the device
    metrics = eval_seg_frames(preds, masks, threshold)                       # one launch, one small D2H
    n = len(metrics)
    return float(losses.sum()) / n, sum(m[0] for m in metrics) / n, sum(m[1] for m in metrics) / n


@torch.no_grad()
def score_volume_lowres(low_segments, mask_dict, frame_ids, obj_list, video_hw, threshold=THRESHOLD, pos_weight=2.0):
    """`score_volume` from the tracker's low-resolution logits (low_segments[frame][obj]: [1,h,w]): up-sampling to
    `video_hw`, thresholded counts and the loss in one pass over the ground truth."""
    lows, masks = [], []
    for f in frame_ids:
        for obj in obj_list:
            p = low_segments[f][obj]
            try:
                m = mask_dict[f][obj].to(dtype=torch.float32, device=p.device).reshape(1, *video_hw)
            except KeyError:
                m = torch.zeros((1,) + tuple(video_hw), dtype=torch.float32, device=p.device)
            lows.append(p.float())
            masks.append(m)
    metrics, losses = score_frames_lowres(torch.stack(lows), torch.stack(masks), threshold, pos_weight)
    n = len(metrics)
    return float(losses.sum()) / n, sum(m[0] for m in metrics) / n, sum(m[1] for m in metrics) / n


@torch.no_grad()
def validation_sam(net, val_loader, prompt="bbox", prompt_freq=2, threshold=THRESHOLD, pos_weight=2.0, shard=False,
                   device=None, fused_scoring=True):
    """Drop-in for the evaluation the reference runs after every epoch (`validation_sam(args, val_loader, epoch, net)`
    with `args.prompt`, `args.prompt_freq`): returns (tot / n_val, (iou / n_val, dice / n_val)).  Like the reference,
    a volume without any annotated object is skipped but still counted in n_val."""
    net.eval()
    device = device or net.device
    packs = list(val_loader)
    n_val = len(packs)
    mine = shard_volumes(n_val) if shard else range(n_val)
    tot, mix = 0.0, [0.0, 0.0]
    for i in mine:
        pack = packs[i]
        imgs = pack["image"]
        if imgs.dim() == 5:
            imgs = imgs.squeeze(0)
        frame_ids = list(range(imgs.size(0)))
        mask_dict = pack["label"]
        obj_list = sorted({o for f in frame_ids for o in mask_dict.get(f, {}).keys()})
        if not obj_list:
            continue
        state = net.val_init_state(imgs_tensor=imgs)
        video_hw = (state["video_height"], state["video_width"])
        fused = (fused_scoring and not getattr(net, "non_overlap_masks", False) and video_hw[1] % 4 == 0
                 and tuple(imgs.shape[2:]) == video_hw)
        state["_ms2_yield_low_res"] = fused          # the tracker yields its low-res logits: no video-res resize / write
        _prompt_volume(net, state, pack, prompt, frame_ids[::prompt_freq], obj_list, tuple(imgs.shape[2:]), device)
        segments = {}
        for f, obj_ids, logits in net.propagate_in_video(state, start_frame_idx=0):
            segments[f] = {o: logits[k] for k, o in enumerate(obj_ids)}
        if fused:
            loss, iou, dice = score_volume_lowres(segments, mask_dict, frame_ids, obj_list, video_hw, threshold, pos_weight)
        else:
            loss, iou, dice = score_volume(segments, mask_dict, frame_ids, obj_list, threshold, pos_weight)
        tot += loss
        mix[0] += iou
        mix[1] += dice
        net.reset_state(state)
    if shard:
        return reduce_validation(tot, mix, len(mine))
    return tot / n_val, (mix[0] / n_val, mix[1] / n_val)
