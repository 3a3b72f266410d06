"""SAM mask decoder (reference modeling/sam/mask_decoder.py:13-317), token-major.

Two-way transformer on [B,Nt,256] tokens x [B,4096,256] image keys; both ConvTranspose2d k2s2
up-scalings are GEMMs (weights re-laid-out to [(dy,dx,co), ci]) followed by a pixel-shuffle kernel
that adds bias + the high-resolution skip feature (+GELU); the 4 hyper-network MLPs feed one
mask-product kernel; the stability fallback counts thresholds on device and selects without a
host synchronisation.  `cell_nums` defaults to None (the fork made it positional and then forgot to
pass it, SURVEY.md §0 finding 1); the repeat_interleave branch of :215-230 is honoured.
"""
import torch
from torch import nn

from ... import ops
from ...runtime import compute_dtype, convT_w_c, p32, par, w_c
from ..sam2_utils import LayerNorm2d, MLP, as_nchw_view, as_nhwc, to_compute


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim, transformer, num_multimask_outputs=3, activation=None,
                 iou_head_depth=3, iou_head_hidden_dim=256, use_high_res_features=False,
                 iou_prediction_use_sigmoid=False, dynamic_multimask_via_stability=False,
                 dynamic_multimask_stability_delta=0.05, dynamic_multimask_stability_thresh=0.98,
                 pred_obj_scores=False, pred_obj_scores_mlp=False, use_multimask_token_for_obj_ptr=False):
        super().__init__()
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.pred_obj_scores = pred_obj_scores
        if self.pred_obj_scores:
            self.obj_score_token = nn.Embedding(1, transformer_dim)
        self.use_multimask_token_for_obj_ptr = use_multimask_token_for_obj_ptr
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            nn.Identity(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            nn.Identity(),
        )
        self.use_high_res_features = use_high_res_features
        if use_high_res_features:
            self.conv_s0 = nn.Conv2d(transformer_dim, transformer_dim // 8, kernel_size=1, stride=1)
            self.conv_s1 = nn.Conv2d(transformer_dim, transformer_dim // 4, kernel_size=1, stride=1)
        self.output_hypernetworks_mlps = nn.ModuleList(
            MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens))
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth,
                                       sigmoid_output=iou_prediction_use_sigmoid)
        if self.pred_obj_scores:
            self.pred_obj_score_head = nn.Linear(transformer_dim, 1)
            if pred_obj_scores_mlp:
                self.pred_obj_score_head = MLP(transformer_dim, transformer_dim, 1, 3)
        self.dynamic_multimask_via_stability = dynamic_multimask_via_stability
        self.dynamic_multimask_stability_delta = dynamic_multimask_stability_delta
        self.dynamic_multimask_stability_thresh = dynamic_multimask_stability_thresh

    # ------------------------------------------------------------------ token-major core
    def _output_tokens(self):
        parts = []
        if self.pred_obj_scores:
            parts.append(self.obj_score_token.weight)
        parts += [self.iou_token.weight, self.mask_tokens.weight]
        from ...runtime import CACHE
        return CACHE.get(tuple(parts), "out_tokens", lambda *ts: torch.cat([t.float() for t in ts], 0).contiguous())

    def predict_tokens(self, src, pos, sparse, feat_s0, feat_s1, tokens=None):
        """src fp32 [B,HW,C] (image embedding + dense prompt); pos fp32 [B,HW,C]; sparse fp32 [B,Ns,C];
        feat_s0 NHWC [B,4h,4w,C/8], feat_s1 NHWC [B,2h,2w,C/4] (or None)
        -> masks fp32 [B,4,4h,4w], iou [B,4], mask tokens fp32 [B,4,C], obj logits [B,1]."""
        cd = compute_dtype()
        B, HW, C = src.shape
        h = w = int(round(HW ** 0.5))
        s = 1 if self.pred_obj_scores else 0
        out_tok = self._output_tokens()
        if tokens is None:                 # `tokens`: output tokens + sparse prompts already in one buffer (prompt encoder)
            tokens = torch.cat((out_tok.unsqueeze(0).expand(B, -1, -1), sparse), dim=1).contiguous()
        hs, keys = self.transformer.forward_tokens(src, pos, tokens)
        Nt = hs.shape[1]
        mask_tokens_out = hs[:, s + 1: s + 1 + self.num_mask_tokens, :]
        dc1, ln1, _, dc2, _ = self.output_upscaling

        def upscale():
            g1 = ops.gemm(to_compute(keys), convT_w_c(dc1.weight), None, out_dtype=torch.float32)     # [B,HW,4*C/4]
            u1 = ops.pixel_shuffle_add(g1, p32(dc1.bias), feat_s1, B, h, w, C // 4)
            u1 = ln1(u1, out_dtype=cd, act=ops.ACT_GELU)                                             # [B,2h,2w,C/4]
            g2 = ops.gemm(u1, convT_w_c(dc2.weight), None, out_dtype=torch.float32)                   # [B,4hw,4*C/8]
            return ops.pixel_shuffle_add(g2, p32(dc2.bias), feat_s0, B, 2 * h, 2 * w, C // 8, act=ops.ACT_GELU)

        hyper = torch.empty((B, self.num_mask_tokens, C // 8), dtype=torch.float32, device=src.device)
        heads = list(self.output_hypernetworks_mlps) + [self.iou_prediction_head]
        obj = None
        grouped_obj = self.pred_obj_scores and isinstance(self.pred_obj_score_head, MLP)
        if grouped_obj:
            heads.append(self.pred_obj_score_head)

        def head_mlps():
            hs2 = to_compute(hs).view(B * Nt, C)        # token rows feed the head MLPs as row-strided GEMM operands
            rows = [hs2[s + 1 + i:: Nt] for i in range(self.num_mask_tokens)] + [hs2[s:: Nt]]     # row-strided [B, C] views
            if grouped_obj:
                rows.append(hs2[0:: Nt])
            # the hyper-network outputs land directly in their rows of `hyper`
            dst = [hyper[:, i, :] for i in range(self.num_mask_tokens)] + [None] * (len(heads) - self.num_mask_tokens)
            outs = self._grouped_mlps(heads, rows, dst) if B <= 64 and len(heads) <= 8 else None
            if outs is None:
                outs = [m(r) for m, r in zip(heads, rows)]
                for i in range(self.num_mask_tokens):
                    hyper[:, i, :] = outs[i]
            return hs2, outs

        # the up-scaling of the image tokens and the head MLPs on the output tokens do not depend on each other
        up, (hs2, outs) = par(upscale, head_mlps)
        masks = ops.hyper_mask(up.view(B, 16 * HW, C // 8), hyper).view(B, self.num_mask_tokens, 4 * h, 4 * w)
        iou_pred = outs[self.num_mask_tokens]
        if self.pred_obj_scores:
            head = self.pred_obj_score_head
            if isinstance(head, MLP):
                obj = outs[self.num_mask_tokens + 1]
            else:
                obj = ops.gemm(to_compute(hs2[0:: Nt].contiguous()), w_c(head.weight), p32(head.bias))
        else:
            obj = 10.0 * iou_pred.new_ones(B, 1)
        return masks, iou_pred, mask_tokens_out, obj

    @staticmethod
    def _grouped_mlps(mlps, rows, dst=None):
        """the head MLPs (4 hyper-networks, IoU, object score) share depth and input width: run them layer by layer
        as ONE grouped small-M launch per layer (3 launches instead of 18).  None if they are not uniform."""
        depth = mlps[0].num_layers
        if any(m.num_layers != depth or m.act != mlps[0].act for m in mlps):
            return None
        cd = compute_dtype()
        xs = list(rows)
        for li in range(depth):
            layers = [m.layers[li] for m in mlps]
            if len({l.weight.shape[1] for l in layers}) != 1 or layers[0].weight.shape[1] % 8:
                return None
            last = li == depth - 1
            acts = [(ops.ACT_SIGMOID if (last and m.sigmoid_output) else ops.ACT_NONE) if last else m.act for m in mlps]
            xs = ops.gemm_grouped(xs, [w_c(l.weight) for l in layers], [p32(l.bias) for l in layers],
                                  out_dtype=torch.float32 if last else cd, acts=acts, outs=dst if last else None)
        return xs

    def select_outputs(self, masks, iou_pred, mask_tokens_out, multimask_output):
        """mask_decoder.py:150-175 without host synchronisation."""
        if multimask_output:
            masks_o, iou_o = masks[:, 1:, :, :], iou_pred[:, 1:]
        elif self.dynamic_multimask_via_stability and not self.training:
            masks_o, iou_o = self._dynamic_multimask_via_stability(masks, iou_pred)
        else:
            masks_o, iou_o = masks[:, 0:1, :, :], iou_pred[:, 0:1]
        if multimask_output and self.use_multimask_token_for_obj_ptr:
            sam_tokens = mask_tokens_out[:, 1:]
        else:
            sam_tokens = mask_tokens_out[:, 0:1]
        return masks_o, iou_o, sam_tokens

    def _dynamic_multimask_via_stability(self, all_mask_logits, all_iou_scores):
        """mask_decoder.py:269-317: threshold counts of the single-mask logits, then stability test, best multimask index
        and its IoU in one launch; the chosen plane is gathered on device (no host synchronisation)."""
        single = all_mask_logits[:, 0:1].contiguous()
        counts = ops.mask_stability_counts(single, self.dynamic_multimask_stability_delta)
        idx, iou_o = ops.stability_select(counts, all_iou_scores.float().contiguous(),
                                          self.dynamic_multimask_stability_thresh)
        return ops.select_plane(all_mask_logits.contiguous(), idx), iou_o

    # ------------------------------------------------------------------ reference signature
    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                multimask_output, repeat_image, cell_nums=None, high_res_features=None):
        B = sparse_prompt_embeddings.size(0)
        if image_embeddings.size(0) != B and cell_nums is not None:
            image_embeddings = torch.repeat_interleave(image_embeddings, cell_nums, dim=0)
        C = image_embeddings.shape[1]
        h, w = image_embeddings.shape[-2:]
        emb = as_nhwc(image_embeddings.float())
        if emb.shape[0] != B:
            emb = emb.expand(B, -1, -1, -1).contiguous()
        dense = dense_prompt_embeddings
        if dense.stride(-1) == 0 and dense.stride(-2) == 0:         # broadcast `no_mask_embed` row
            vec = dense[:, :, 0, 0].float()
            if vec.stride(0) == 0 or vec.shape[0] == 1:
                src = ops.add_rowvec(emb, vec[0].contiguous())
            else:
                src = ops.axpby(emb, 1.0, as_nhwc(dense.float().contiguous()), 1.0)
        else:
            d = as_nhwc(dense.float())
            if d.shape[0] != B:
                d = d.expand(B, -1, -1, -1).contiguous()
            src = ops.axpby(emb, 1.0, d, 1.0)
        pos = as_nhwc(image_pe.float()).reshape(1, h * w, C)
        pos = pos.expand(B, -1, -1).contiguous() if B > 1 else pos
        f0 = f1 = None
        if self.use_high_res_features and high_res_features is not None:
            f0, f1 = (as_nhwc(f.float()) for f in high_res_features)
            if f0.shape[0] != B:
                f0, f1 = f0.expand(B, -1, -1, -1).contiguous(), f1.expand(B, -1, -1, -1).contiguous()
        tokens = getattr(sparse_prompt_embeddings, "_ms2_tokens", None)
        n_out = self.num_mask_tokens + 1 + (1 if self.pred_obj_scores else 0)
        if tokens is not None and not (tokens.dtype == torch.float32 and tokens.is_contiguous() and tokens.shape[0] == B
                                       and tokens.shape[1] == n_out + sparse_prompt_embeddings.shape[1]):
            tokens = None
        masks, iou_pred, mask_tokens_out, obj = self.predict_tokens(
            src.view(B, h * w, C), pos, None if tokens is not None else sparse_prompt_embeddings.float().contiguous(),
            f0, f1, tokens=tokens)
        masks_o, iou_o, sam_tokens = self.select_outputs(masks, iou_pred, mask_tokens_out, multimask_output)
        return masks_o, iou_o, sam_tokens, obj
