"""SAM2AutomaticMaskGenerator (reference `sam2_train/automatic_mask_generator.py:36-434`).

Same constructor arguments, `generate(image)` records and filtering rules as the reference.  What differs is where the
candidates live: the reference carries every candidate of a point batch as full-resolution tensors through
`calculate_stability_score`, a binarisation, `batched_mask_to_box`, `uncrop_masks` and a per-mask `nonzero`; here one
kernel pass over the logits (`ops.mask_stats`) returns the three threshold counts and the box of every candidate, the
filters run on those few integers on the host, and only the survivors are binarised, un-cropped and transposed on the
GPU (`ops.mask_binarize_t`) and their run boundaries compacted there (`ops.rle_transitions`), so a survivor costs a few
KB of D2H traffic instead of its H*W mask.  Box NMS runs on the host over the surviving boxes.
"""
import numpy as np
import torch

from .sam2_image_predictor import SAM2ImagePredictor
from .utils import amg


class _Candidates:
    """Column store for the surviving candidates of one crop / image (host side; the reference's `MaskData`)."""
    FIELDS = ("rles", "boxes", "iou_preds", "points", "stability_score", "crop_boxes")

    def __init__(self):
        self.rles = []
        self.boxes = np.zeros((0, 4), np.int64)
        self.iou_preds = np.zeros((0,), np.float32)
        self.points = np.zeros((0, 2), np.float64)
        self.stability_score = np.zeros((0,), np.float32)
        self.crop_boxes = np.zeros((0, 4), np.int64)

    def __len__(self):
        return len(self.rles)

    def extend(self, other):
        self.rles += other.rles
        for f in self.FIELDS[1:]:
            setattr(self, f, np.concatenate([getattr(self, f), getattr(other, f)], axis=0))

    def take(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        self.rles = [self.rles[i] for i in idx]
        for f in self.FIELDS[1:]:
            setattr(self, f, getattr(self, f)[idx])


class SAM2AutomaticMaskGenerator:
    def __init__(self, model, points_per_side=32, points_per_batch=64, pred_iou_thresh=0.8, stability_score_thresh=0.95,
                 stability_score_offset=1.0, mask_threshold=0.0, box_nms_thresh=0.7, crop_n_layers=0,
                 crop_nms_thresh=0.7, crop_overlap_ratio=512 / 1500, crop_n_points_downscale_factor=1, point_grids=None,
                 min_mask_region_area=0, output_mode="binary_mask", use_m2m=False, multimask_output=True):
        assert (points_per_side is None) != (point_grids is None), \
            "Exactly one of points_per_side or point_grid must be provided."
        if points_per_side is not None:
            self.point_grids = amg.build_all_layer_point_grids(points_per_side, crop_n_layers,
                                                               crop_n_points_downscale_factor)
        else:
            self.point_grids = point_grids
        assert output_mode in ["binary_mask", "uncompressed_rle", "coco_rle"], f"Unknown output_mode {output_mode}."
        if output_mode == "coco_rle":
            from pycocotools import mask as mask_utils  # type: ignore  # noqa: F401
        self.predictor = SAM2ImagePredictor(model, max_hole_area=min_mask_region_area,
                                            max_sprinkle_area=min_mask_region_area)
        self.points_per_batch = points_per_batch
        self.pred_iou_thresh = pred_iou_thresh
        self.stability_score_thresh = stability_score_thresh
        self.stability_score_offset = stability_score_offset
        self.mask_threshold = mask_threshold
        self.box_nms_thresh = box_nms_thresh
        self.crop_n_layers = crop_n_layers
        self.crop_nms_thresh = crop_nms_thresh
        self.crop_overlap_ratio = crop_overlap_ratio
        self.crop_n_points_downscale_factor = crop_n_points_downscale_factor
        self.min_mask_region_area = min_mask_region_area
        self.output_mode = output_mode
        self.use_m2m = use_m2m
        self.multimask_output = multimask_output

    # ------------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def generate(self, image):
        """image HWC uint8 -> list of records {segmentation, area, bbox (XYWH), predicted_iou, point_coords,
        stability_score, crop_box (XYWH)} (automatic_mask_generator.py:152-211)."""
        data = self._generate_masks(image)
        if self.output_mode == "coco_rle":
            segs = [amg.coco_encode_rle(r) for r in data.rles]
        elif self.output_mode == "binary_mask":
            segs = [amg.rle_to_mask(r) for r in data.rles]
        else:
            segs = data.rles
        return [{"segmentation": segs[i],
                 "area": amg.area_from_rle(data.rles[i]),
                 "bbox": amg.box_xyxy_to_xywh(data.boxes[i]).tolist(),
                 "predicted_iou": float(data.iou_preds[i]),
                 "point_coords": [data.points[i].tolist()],
                 "stability_score": float(data.stability_score[i]),
                 "crop_box": amg.box_xyxy_to_xywh(data.crop_boxes[i]).tolist()} for i in range(len(data))]

    # ------------------------------------------------------------------------------------------ crops
    def _generate_masks(self, image):
        """automatic_mask_generator.py:213-240."""
        orig_size = image.shape[:2]
        crop_boxes, layer_idxs = amg.generate_crop_boxes(orig_size, self.crop_n_layers, self.crop_overlap_ratio)
        data = _Candidates()
        for crop_box, layer_idx in zip(crop_boxes, layer_idxs):
            data.extend(self._process_crop(image, crop_box, layer_idx, orig_size))
        if len(crop_boxes) > 1 and len(data):
            cb = data.crop_boxes.astype(np.float32)
            scores = 1 / ((cb[:, 2] - cb[:, 0]) * (cb[:, 3] - cb[:, 1]))          # prefer masks from smaller crops
            data.take(amg.nms(data.boxes, scores, self.crop_nms_thresh))
        return data

    def _process_crop(self, image, crop_box, crop_layer_idx, orig_size):
        """automatic_mask_generator.py:242-285."""
        x0, y0, x1, y1 = crop_box
        cropped = image[y0:y1, x0:x1, :]
        crop_hw = cropped.shape[:2]
        self.predictor.set_image(cropped)
        points = self.point_grids[crop_layer_idx] * np.array(crop_hw)[None, ::-1]
        data = _Candidates()
        for (pts,) in amg.batch_iterator(self.points_per_batch, points):
            data.extend(self._process_batch(pts, crop_hw, crop_box, orig_size, normalize=True))
        self.predictor.reset_predictor()
        if len(data):
            data.take(amg.nms(data.boxes, data.iou_preds, self.box_nms_thresh))
        data.boxes = data.boxes + np.array([[x0, y0, x0, y0]])
        data.points = data.points + np.array([[x0, y0]])
        data.crop_boxes = np.tile(np.asarray(crop_box, np.int64)[None], (len(data), 1))
        return data

    # ------------------------------------------------------------------------------------------ one batch of points
    def _predict_points(self, points_np, im_size, normalize, mask_input=None, multimask=True):
        pts = torch.as_tensor(points_np, device=self.predictor.device)
        in_points = self.predictor._transforms.transform_coords(pts, normalize=normalize, orig_hw=im_size)
        in_labels = torch.ones(in_points.shape[0], dtype=torch.int, device=in_points.device)
        return self.predictor._predict(in_points[:, None, :], in_labels[:, None], mask_input=mask_input,
                                       multimask_output=multimask, return_logits=True)

    def _process_batch(self, points, im_size, crop_box, orig_size, normalize=False):
        """automatic_mask_generator.py:287-372: predict, filter by predicted IoU and stability, drop boxes on crop edges,
        run-length encode the survivors."""
        orig_h, orig_w = orig_size
        masks, iou_preds, low_res = self._predict_points(points, im_size, normalize, multimask=self.multimask_output)
        per_point = masks.shape[1]
        masks = masks.flatten(0, 1)                                              # [N,H,W] logits, stay on the GPU
        ious = iou_preds.flatten(0, 1).float().cpu().numpy()
        pts = np.repeat(np.asarray(points, dtype=np.float64), per_point, axis=0)
        keep = np.arange(len(ious))
        if self.use_m2m:                                                         # one refinement step from the low-res logits
            low = low_res.flatten(0, 1)
            new_masks, new_ious = [], []
            for p, l in amg.batch_iterator(self.points_per_batch, pts, low):
                m2, i2, _ = self._predict_points(p, im_size, normalize, mask_input=l[:, None, :], multimask=False)
                new_masks.append(m2)
                new_ious.append(i2)
            masks = torch.cat(new_masks, dim=0).squeeze(1)
            ious = torch.cat(new_ious, dim=0).squeeze(1).float().cpu().numpy()
        stats = amg.mask_stats(masks, self.mask_threshold, self.stability_score_offset)
        stability = amg.stability_from_stats(stats)
        if self.pred_iou_thresh > 0.0:
            keep = keep[ious[keep] > self.pred_iou_thresh]
        if self.stability_score_thresh > 0.0:
            keep = keep[stability[keep] >= self.stability_score_thresh]
        boxes = amg.boxes_from_stats(stats)
        near_edge = amg.is_box_near_crop_edge(boxes[keep], crop_box, [0, 0, orig_w, orig_h])
        keep = keep[~near_edge]
        out = _Candidates()
        if len(keep):
            sel = torch.as_tensor(keep, dtype=torch.int32, device=masks.device)
            out.rles = amg.rles_from_device(self._binarize_survivors(masks, sel, crop_box, orig_h, orig_w))
        out.boxes, out.iou_preds, out.points = boxes[keep], ious[keep], pts[keep]
        out.stability_score = stability[keep]
        out.crop_boxes = np.zeros((len(keep), 4), np.int64)
        return out

    def _binarize_survivors(self, masks, sel, crop_box, orig_h, orig_w):
        """(masks[sel] > mask_threshold) un-cropped to the original image and transposed, uint8 [K,W,H] on the GPU."""
        from . import ops
        m = masks if masks.dtype == torch.float32 else masks.float()
        return ops.mask_binarize_t(m.contiguous(), sel, self.mask_threshold, (orig_h, orig_w),
                                   (crop_box[0], crop_box[1]))

    # ------------------------------------------------------------------------------------------ optional clean-up
    @staticmethod
    def postprocess_small_regions(mask_data, min_area, nms_thresh):
        """automatic_mask_generator.py:374-420: remove small islands / holes of every mask, then re-run box NMS
        preferring masks that did not change.  Works on (and returns) the candidate store of `_generate_masks`."""
        if len(mask_data) == 0:
            return mask_data
        new_masks, scores = [], []
        for rle in mask_data.rles:
            mask = amg.rle_to_mask(rle)
            mask, changed_h = amg.remove_small_regions(mask, min_area, mode="holes")
            mask, changed_i = amg.remove_small_regions(mask, min_area, mode="islands")
            new_masks.append(mask)
            scores.append(float(not (changed_h or changed_i)))
        boxes = np.zeros((len(new_masks), 4), np.int64)
        for i, m in enumerate(new_masks):
            ys, xs = np.nonzero(m)
            if len(xs):
                boxes[i] = [xs.min(), ys.min(), xs.max(), ys.max()]
        keep = amg.nms(boxes, scores, nms_thresh)
        for i in keep:
            if scores[i] == 0.0:
                mask_data.rles[i] = amg.rle_from_transposed(np.ascontiguousarray(new_masks[i].T).astype(np.uint8))
                mask_data.boxes[i] = boxes[i]
        mask_data.take(keep)
        return mask_data
