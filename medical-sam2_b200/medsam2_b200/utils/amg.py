"""Helpers of the automatic mask generator (reference `sam2_train/utils/amg.py`), re-cut for the B200 path.

The reference keeps every candidate as full-resolution tensors and filters them with a dozen whole-plane torch
reductions (`calculate_stability_score` :158-180, `batched_mask_to_box` :296-348, `mask_to_rle_pytorch` :107-134).  Here
the per-plane work is two kernels (`ops.mask_stats`: three threshold counts + the bounding box in ONE pass;
`ops.mask_binarize_t`: binarise + un-crop + transpose of the survivors only) and everything that remains is arithmetic
on a few integers per candidate, done on the host in numpy.  Function names and results follow the reference.
"""
import math
from itertools import product

import numpy as np
import torch

from .. import ops


# ------------------------------------------------------------------------------------------ grids and crops (host)
def build_point_grid(n_per_side):
    """utils/amg.py:183-191: n x n points at the cell centres of the unit square, x fastest."""
    off = 1 / (2 * n_per_side)
    side = np.linspace(off, 1 - off, n_per_side)
    xs, ys = np.meshgrid(side, side, indexing="xy")
    return np.stack([xs, ys], axis=-1).reshape(-1, 2)


def build_all_layer_point_grids(n_per_side, n_layers, scale_per_layer):
    """utils/amg.py:194-203."""
    return [build_point_grid(int(n_per_side / (scale_per_layer ** i))) for i in range(n_layers + 1)]


def generate_crop_boxes(im_size, n_layers, overlap_ratio):
    """utils/amg.py:206-243: the full image, then (2^i)^2 overlapping crops per extra layer; XYXY boxes + layer index."""
    im_h, im_w = im_size
    short = min(im_h, im_w)
    boxes, layers = [[0, 0, im_w, im_h]], [0]
    for layer in range(n_layers):
        n = 2 ** (layer + 1)
        overlap = int(overlap_ratio * short * (2 / n))
        cw = int(math.ceil((overlap * (n - 1) + im_w) / n))
        ch = int(math.ceil((overlap * (n - 1) + im_h) / n))
        for x0, y0 in product([int((cw - overlap) * i) for i in range(n)], [int((ch - overlap) * i) for i in range(n)]):
            boxes.append([x0, y0, min(x0 + cw, im_w), min(y0 + ch, im_h)])
            layers.append(layer + 1)
    return boxes, layers


def batch_iterator(batch_size, *args):
    """utils/amg.py:97-104."""
    assert len(args) > 0 and all(len(a) == len(args[0]) for a in args), \
        "Batched iteration must have inputs of all the same size."
    for b in range(0, len(args[0]), batch_size):
        yield [a[b:b + batch_size] for a in args]


# ------------------------------------------------------------------------------------------ per-candidate statistics
def mask_stats(masks, mask_threshold, threshold_offset):
    """masks fp32 logits [N,H,W] on the GPU -> numpy int64 [N,7] (see `ops.mask_stats`); one launch, one small D2H."""
    m = masks if masks.dtype == torch.float32 else masks.float()
    return ops.mask_stats(m.contiguous(), mask_threshold, threshold_offset).cpu().numpy().astype(np.int64)


def stability_from_stats(stats):
    """`calculate_stability_score` (utils/amg.py:158-180) from the counts: int32 / int32 true division in torch gives
    fp32 (0/0 -> nan, n/0 -> inf like the reference)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return (stats[:, 0].astype(np.float32) / stats[:, 1].astype(np.float32)).astype(np.float32)


def boxes_from_stats(stats):
    """`batched_mask_to_box` (utils/amg.py:296-348): XYXY with inclusive max indices, [0,0,0,0] for an empty mask."""
    boxes = stats[:, 3:7].copy()
    boxes[stats[:, 2] == 0] = 0
    return boxes


def calculate_stability_score(masks, mask_threshold, threshold_offset):
    """Drop-in for utils/amg.py:158-180: masks [..., H, W] logits on the GPU -> fp32 tensor [...] on the same device."""
    lead = masks.shape[:-2]
    st = mask_stats(masks.reshape(-1, *masks.shape[-2:]), mask_threshold, threshold_offset)
    return torch.from_numpy(stability_from_stats(st)).to(masks.device).reshape(lead)


def batched_mask_to_box(masks):
    """Drop-in for utils/amg.py:296-348: boolean (or 0/1) masks [..., H, W] on the GPU -> int64 boxes [..., 4]."""
    if masks.numel() == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    lead = masks.shape[:-2]
    st = mask_stats(masks.reshape(-1, *masks.shape[-2:]).float(), 0.5, 0.0)
    return torch.from_numpy(boxes_from_stats(st)).to(masks.device).reshape(*lead, 4)


def is_box_near_crop_edge(boxes, crop_box, orig_box, atol=20.0):
    """utils/amg.py:80-91 on numpy boxes [N,4] (crop coordinates): touches a crop edge that is not an image edge."""
    x0, y0 = crop_box[0], crop_box[1]
    b = boxes.astype(np.float32) + np.array([x0, y0, x0, y0], np.float32)
    near_crop = np.abs(b - np.asarray(crop_box, np.float32)[None]) <= atol
    near_image = np.abs(b - np.asarray(orig_box, np.float32)[None]) <= atol
    return np.any(near_crop & ~near_image, axis=1)


# ------------------------------------------------------------------------------------------ run-length encoding
def rle_from_transposed(mt):
    """mt uint8 [W,H] = the mask in column-major order (what `ops.mask_binarize_t` returns, on the host) -> the
    uncompressed RLE of utils/amg.py:107-134: alternating run lengths starting with the zeros."""
    w, h = mt.shape
    flat = mt.reshape(-1)
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    edges = np.concatenate([[0], change, [h * w]])
    counts = ([] if flat[0] == 0 else [0]) + np.diff(edges).tolist()
    return {"size": [h, w], "counts": counts}


def rles_from_device(mt, cap=4096):
    """mt uint8 [K,W,H] on the GPU (column-major masks from `ops.mask_binarize_t`) -> list of uncompressed RLE dicts.
    The change positions are found and compacted on the device (`ops.rle_transitions`); the host receives cap positions
    per mask (16 KB instead of the W*H-byte mask) and turns them into run lengths; a mask with more transitions than
    `cap` triggers one re-run with the exact capacity."""
    K, w, h = mt.shape
    if K == 0:
        return []
    flat = mt.reshape(K, -1)
    pos, cnt = ops.rle_transitions(flat, cap)
    cnt_h = cnt.cpu().numpy()
    if int(cnt_h.max()) > cap:
        pos, _ = ops.rle_transitions(flat, int(cnt_h.max()))
    pos_h, first = pos.cpu().numpy(), flat[:, 0].cpu().numpy()
    out = []
    for k in range(K):
        edges = np.concatenate([[0], pos_h[k, :cnt_h[k]], [h * w]])
        out.append({"size": [h, w], "counts": ([] if first[k] == 0 else [0]) + np.diff(edges).tolist()})
    return out


def mask_to_rle_pytorch(tensor):
    """Drop-in for utils/amg.py:107-134: boolean masks [b,h,w] on the GPU -> list of uncompressed RLE dicts."""
    b, h, w = tensor.shape
    sel = torch.arange(b, dtype=torch.int32, device=tensor.device)
    return rles_from_device(ops.mask_binarize_t(tensor.float().contiguous(), sel, 0.5, (h, w), (0, 0)))


def rle_to_mask(rle):
    """utils/amg.py:137-150: uncompressed RLE -> boolean [h,w] numpy mask."""
    h, w = rle["size"]
    counts = np.asarray(rle["counts"], dtype=np.int64)
    vals = (np.arange(len(counts)) & 1).astype(bool)
    return np.repeat(vals, counts).reshape(w, h).transpose()


def area_from_rle(rle):
    """utils/amg.py:153-154."""
    return sum(rle["counts"][1::2])


def coco_encode_rle(uncompressed_rle):
    """utils/amg.py:285-293 (needs pycocotools, like the reference)."""
    from pycocotools import mask as mask_utils  # type: ignore
    h, w = uncompressed_rle["size"]
    rle = mask_utils.frPyObjects(uncompressed_rle, h, w)
    rle["counts"] = rle["counts"].decode("utf-8")
    return rle


def box_xyxy_to_xywh(box):
    """utils/amg.py:88-93."""
    b = np.array(box).copy()
    b[2] -= b[0]
    b[3] -= b[1]
    return b


# ------------------------------------------------------------------------------------------ non-maximum suppression
def nms(boxes, scores, iou_threshold):
    """`torchvision.ops.batched_nms` with a single category, as automatic_mask_generator.py:226-232,263-269 call it, on a
    few hundred host-side boxes: visit by descending score, keep a box unless IoU with a kept one exceeds the threshold.
    Returns the kept indices in visiting order (torchvision's order)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    scores = np.asarray(scores, np.float32).reshape(-1)
    order = np.argsort(-scores, kind="stable")
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    keep, dead = [], np.zeros(len(boxes), bool)
    for i in order:
        if dead[i]:
            continue
        keep.append(int(i))
        lt = np.maximum(boxes[i, :2], boxes[:, :2])
        rb = np.minimum(boxes[i, 2:], boxes[:, 2:])
        wh = np.clip(rb - lt, 0, None)
        inter = wh[:, 0] * wh[:, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / (area[i] + area - inter)
        dead |= iou > iou_threshold
    return np.asarray(keep, dtype=np.int64)


def remove_small_regions(mask, area_thresh, mode):
    """utils/amg.py:246-282 with scipy's 8-connected labelling in place of OpenCV's (a host-side helper of the optional
    `postprocess_small_regions`; `generate()` itself never calls it)."""
    from scipy import ndimage
    assert mode in ["holes", "islands"]
    holes = mode == "holes"
    work = (holes ^ mask).astype(np.uint8)
    regions, n = ndimage.label(work, structure=np.ones((3, 3), np.uint8))
    sizes = np.bincount(regions.reshape(-1), minlength=n + 1)[1:]
    small = [i + 1 for i, s in enumerate(sizes) if s < area_thresh]
    if not small:
        return mask, False
    fill = [0] + small
    if not holes:
        fill = [i for i in range(n + 1) if i not in fill]
        if not fill:
            fill = [int(np.argmax(sizes)) + 1]
    return np.isin(regions, fill), True
