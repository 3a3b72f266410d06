"""SAM2Transforms (reference utils/transforms.py:13-99) on the native kernels: uint8 HWC ->
normalised fp32 NCHW in one kernel (antialiased bilinear resize first when the image is not already
at the model resolution), coordinate scaling, and mask post-processing with the CC kernel."""
import numpy as np
import torch
from torch import nn

from .. import ops


class SAM2Transforms(nn.Module):
    def __init__(self, resolution, mask_threshold, max_hole_area=0.0, max_sprinkle_area=0.0):
        super().__init__()
        self.resolution = resolution
        self.mask_threshold = mask_threshold
        self.max_hole_area = max_hole_area
        self.max_sprinkle_area = max_sprinkle_area
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]

    def _one(self, img, device):
        if not isinstance(img, np.ndarray):
            img = np.array(img)
        x = torch.from_numpy(np.ascontiguousarray(img)).to(device)
        assert x.dtype == torch.uint8 and x.dim() == 3 and x.shape[-1] == 3, "expected an HWC uint8 RGB image"
        H, W = x.shape[:2]
        if (H, W) == (self.resolution, self.resolution):
            return ops.normalize_image(x[None], nhwc=True)[0]
        planes = x.permute(2, 0, 1).float().contiguous()            # [3,H,W] in 0..255
        planes = ops.resize_bilinear(planes, (self.resolution, self.resolution), antialias=True)
        return ops.normalize_image(planes[None].contiguous())[0]

    def __call__(self, x, device="cuda"):
        return self._one(x, device)

    def forward_batch(self, img_list, device="cuda"):
        return torch.stack([self._one(img, device) for img in img_list], dim=0)

    def transform_coords(self, coords, normalize=False, orig_hw=None):
        if normalize:
            assert orig_hw is not None
            h, w = orig_hw
            coords = coords.clone()
            coords[..., 0] = coords[..., 0] / w
            coords[..., 1] = coords[..., 1] / h
        return coords * self.resolution

    def transform_boxes(self, boxes, normalize=False, orig_hw=None):
        return self.transform_coords(boxes.reshape(-1, 2, 2), normalize, orig_hw)

    def postprocess_masks(self, masks, orig_hw):
        """utils/transforms.py:74-99."""
        from .misc import get_connected_components
        masks = masks.float().contiguous()
        mask_flat = None
        if self.max_hole_area > 0:
            mask_flat = masks.flatten(0, 1).unsqueeze(1)
            labels, areas = get_connected_components(mask_flat <= self.mask_threshold)
            is_hole = ((labels > 0) & (areas <= self.max_hole_area)).reshape_as(masks)
            masks = torch.where(is_hole, self.mask_threshold + 10.0, masks)
        if self.max_sprinkle_area > 0:
            if mask_flat is None:
                mask_flat = masks.flatten(0, 1).unsqueeze(1)
            labels, areas = get_connected_components(mask_flat > self.mask_threshold)
            is_hole = ((labels > 0) & (areas <= self.max_sprinkle_area)).reshape_as(masks)
            masks = torch.where(is_hole, self.mask_threshold - 10.0, masks)
        return ops.resize_bilinear(masks.contiguous(), tuple(orig_hw))
