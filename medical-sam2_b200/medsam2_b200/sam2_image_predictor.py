"""SAM2ImagePredictor with the reference's interface (sam2_image_predictor.py:20-446)."""
import logging

import numpy as np
import torch

from . import ops
from .modeling.sam2_utils import seq_to_tokens
from .runtime import p32
from .utils.transforms import SAM2Transforms



def _to_numpy(tensors):
    """device tensors -> numpy arrays (the reference does `.float().detach().cpu().numpy()` per tensor: a pageable,
    synchronous copy each - 12.6 MB of logits per 1024^2 image).  All copies go to pinned buffers of torch's caching
    host allocator without blocking, followed by ONE stream synchronisation."""
    outs = []
    cuda = False
    for t in tensors:
        t = t.detach().float()
        if t.is_cuda:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            cuda = True
        else:
            h = t
        outs.append(h)
    if cuda:
        torch.cuda.current_stream().synchronize()
    return [h.numpy() for h in outs]


class SAM2ImagePredictor:
    def __init__(self, sam_model, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0):
        self.model = sam_model
        self._transforms = SAM2Transforms(resolution=self.model.image_size, mask_threshold=mask_threshold,
                                          max_hole_area=max_hole_area, max_sprinkle_area=max_sprinkle_area)
        self.mask_threshold = mask_threshold
        s = self.model.image_size
        self._bb_feat_sizes = [(s // 4, s // 4), (s // 8, s // 8), (s // 16, s // 16)]
        self.reset_predictor()

    # ------------------------------------------------------------------ embedding
    def _embed(self, batch):
        m = self.model
        backbone_out = m.forward_image(batch)
        _, vision_feats, _, _ = m._prepare_backbone_features(backbone_out)
        B = batch.shape[0]
        feats = [f.permute(1, 2, 0).reshape(B, -1, *fs) for f, fs in zip(vision_feats, self._bb_feat_sizes)]
        if m.directly_add_no_mem_embed:      # sam2_image_predictor.py:143-144, as one row-broadcast kernel
            h, w = self._bb_feat_sizes[-1]
            tok = seq_to_tokens(vision_feats[-1].float())
            tok = ops.add_rowvec(tok, p32(m.no_mem_embed).view(-1))
            feats[-1] = tok.view(B, h, w, -1).permute(0, 3, 1, 2)
        self._features = {"image_embed": feats[-1], "high_res_feats": feats[:-1]}
        self._is_image_set = True

    @torch.no_grad()
    def set_image(self, image):
        self.reset_predictor()
        if isinstance(image, np.ndarray):
            logging.info("For numpy array image, we assume (HxWxC) format")
            self._orig_hw = [image.shape[:2]]
        elif hasattr(image, "size") and not isinstance(image, torch.Tensor):
            w, h = image.size
            self._orig_hw = [(h, w)]
        else:
            raise NotImplementedError("Image format not supported")
        x = self._transforms(image, device=self.device)[None]
        assert x.dim() == 4 and x.shape[1] == 3, f"input_image must be of size 1x3xHxW, got {x.shape}"
        self._embed(x)

    @torch.no_grad()
    def set_image_batch(self, image_list):
        self.reset_predictor()
        assert isinstance(image_list, list)
        self._orig_hw = []
        for image in image_list:
            assert isinstance(image, np.ndarray), \
                "Images are expected to be an np.ndarray in RGB format, and of shape  HWC"
            self._orig_hw.append(image.shape[:2])
        batch = self._transforms.forward_batch(image_list, device=self.device)
        assert batch.dim() == 4 and batch.shape[1] == 3, f"img_batch must be of size Bx3xHxW, got {batch.shape}"
        self._embed(batch)
        self._is_batch = True

    # ------------------------------------------------------------------ prediction
    def predict_batch(self, point_coords_batch=None, point_labels_batch=None, box_batch=None, mask_input_batch=None,
                      multimask_output=True, return_logits=False, normalize_coords=True):
        assert self._is_batch, "This function should only be used when in batched mode"
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image_batch(...) before mask prediction.")
        n = len(self._features["image_embed"])
        all_masks, all_ious, all_low = [], [], []
        pick = lambda b, i: b[i] if b is not None else None
        for i in range(n):
            mask_input, coords, labels, box = self._prep_prompts(
                pick(point_coords_batch, i), pick(point_labels_batch, i), pick(box_batch, i),
                pick(mask_input_batch, i), normalize_coords, img_idx=i)
            masks, ious, low = self._predict(coords, labels, box, mask_input, multimask_output,
                                             return_logits=return_logits, img_idx=i)
            all_masks.append(masks.squeeze(0))
            all_ious.append(ious.squeeze(0))
            all_low.append(low.squeeze(0))
        host = _to_numpy(all_masks + all_ious + all_low)          # one synchronisation for the whole batch
        return host[:n], host[n: 2 * n], host[2 * n:]

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False, normalize_coords=True):
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        mask_input, coords, labels, box = self._prep_prompts(point_coords, point_labels, box, mask_input,
                                                             normalize_coords)
        masks, ious, low = self._predict(coords, labels, box, mask_input, multimask_output, return_logits=return_logits)
        return tuple(_to_numpy([masks.squeeze(0), ious.squeeze(0), low.squeeze(0)]))

    def _prep_prompts(self, point_coords, point_labels, box, mask_logits, normalize_coords, img_idx=-1):
        coords = labels = ubox = mask_input = None
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied if point_coords is supplied."
            pc = torch.as_tensor(point_coords, dtype=torch.float, device=self.device)
            coords = self._transforms.transform_coords(pc, normalize=normalize_coords, orig_hw=self._orig_hw[img_idx])
            labels = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)
            if coords.dim() == 2:
                coords, labels = coords[None, ...], labels[None, ...]
        if box is not None:
            b = torch.as_tensor(box, dtype=torch.float, device=self.device)
            ubox = self._transforms.transform_boxes(b, normalize=normalize_coords, orig_hw=self._orig_hw[img_idx])
        if mask_logits is not None:
            mask_input = torch.as_tensor(mask_logits, dtype=torch.float, device=self.device)
            if mask_input.dim() == 3:
                mask_input = mask_input[None, :, :, :]
        return mask_input, coords, labels, ubox

    @torch.no_grad()
    def _predict(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                 return_logits=False, img_idx=-1):
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        concat = (point_coords, point_labels) if point_coords is not None else None
        if boxes is not None:
            bc = boxes.reshape(-1, 2, 2)
            bl = torch.tensor([[2, 3]], dtype=torch.int, device=boxes.device).repeat(boxes.size(0), 1)
            if concat is not None:
                concat = (torch.cat([bc, concat[0]], dim=1), torch.cat([bl, concat[1]], dim=1))
            else:
                concat = (bc, bl)
        m = self.model
        sparse, dense = m.sam_prompt_encoder(points=concat, boxes=None, masks=mask_input)
        batched_mode = concat is not None and concat[0].shape[0] > 1
        hr = [lvl[img_idx].unsqueeze(0) for lvl in self._features["high_res_feats"]]
        low, ious, _, _ = m.sam_mask_decoder(
            image_embeddings=self._features["image_embed"][img_idx].unsqueeze(0),
            image_pe=m.sam_prompt_encoder.get_dense_pe(), sparse_prompt_embeddings=sparse,
            dense_prompt_embeddings=dense, multimask_output=multimask_output, repeat_image=batched_mode,
            high_res_features=hr)
        masks = self._transforms.postprocess_masks(low, self._orig_hw[img_idx])
        low = torch.clamp(low, -32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks, ious, low

    def get_image_embedding(self):
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        assert self._features is not None, "Features must exist if an image has been set."
        return self._features["image_embed"]

    @property
    def device(self):
        return self.model.device

    def reset_predictor(self):
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False
