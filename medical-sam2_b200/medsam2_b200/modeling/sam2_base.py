"""SAM2Base: per-frame glue of the hot path (reference modeling/sam2_base.py:22-830).

Same constructor arguments, sub-module attribute names, learned tokens and method signatures as the
reference, so `state_dict`s load strictly and `func_2d` / `func_3d` style callers work unchanged;
`image_size` follows the config (the fork hard-codes 256, SURVEY.md §0 finding 1).  Internally every
tensor is token-major and all arithmetic runs in the kernels of `medsam2_b200.ops`.
"""
import torch
from torch import nn
from torch.nn.init import trunc_normal_

from .. import ops
from ..runtime import CACHE, compute_dtype, conv_w_c, p32, par, w_c
from .sam.mask_decoder import MaskDecoder
from .sam.prompt_encoder import PromptEncoder
from .sam.transformer import TwoWayTransformer
from .sam2_utils import (MLP, Linear, as_nchw_view, as_nhwc, select_closest_cond_frames, seq_to_tokens,
                         to_compute)

NO_OBJ_SCORE = -1024.0


class SAM2Base(nn.Module):
    def __init__(self, image_encoder, memory_attention, memory_encoder, num_maskmem=7, image_size=512,
                 backbone_stride=16, sigmoid_scale_for_mem_enc=1.0, sigmoid_bias_for_mem_enc=0.0,
                 binarize_mask_from_pts_for_mem_enc=False, use_mask_input_as_output_without_sam=False,
                 max_cond_frames_in_attn=-1, directly_add_no_mem_embed=False, use_high_res_features_in_sam=False,
                 multimask_output_in_sam=False, multimask_min_pt_num=1, multimask_max_pt_num=1,
                 multimask_output_for_tracking=False, use_multimask_token_for_obj_ptr=False,
                 iou_prediction_use_sigmoid=False, memory_temporal_stride_for_eval=1,
                 add_all_frames_to_correct_as_cond=False, non_overlap_masks_for_mem_enc=False,
                 use_obj_ptrs_in_encoder=False, max_obj_ptrs_in_encoder=16, add_tpos_enc_to_obj_ptrs=True,
                 proj_tpos_enc_in_obj_ptrs=False, only_obj_ptrs_in_the_past_for_eval=False,
                 pred_obj_scores=False, pred_obj_scores_mlp=False, fixed_no_obj_ptr=False, soft_no_obj_ptr=False,
                 use_mlp_for_obj_ptr_proj=False, sam_mask_decoder_extra_args=None, compile_image_encoder=False):
        super().__init__()
        self.image_encoder = image_encoder
        self.use_high_res_features_in_sam = use_high_res_features_in_sam
        self.num_feature_levels = 3 if use_high_res_features_in_sam else 1
        self.use_obj_ptrs_in_encoder = use_obj_ptrs_in_encoder
        self.max_obj_ptrs_in_encoder = max_obj_ptrs_in_encoder
        if use_obj_ptrs_in_encoder:
            self.mask_downsample = nn.Conv2d(1, 1, kernel_size=4, stride=4)
        self.add_tpos_enc_to_obj_ptrs = add_tpos_enc_to_obj_ptrs
        if proj_tpos_enc_in_obj_ptrs:
            assert add_tpos_enc_to_obj_ptrs
        self.proj_tpos_enc_in_obj_ptrs = proj_tpos_enc_in_obj_ptrs
        self.only_obj_ptrs_in_the_past_for_eval = only_obj_ptrs_in_the_past_for_eval
        self.memory_attention = memory_attention
        self.hidden_dim = memory_attention.d_model
        self.memory_encoder = memory_encoder
        self.mem_dim = self.hidden_dim
        if hasattr(self.memory_encoder, "out_proj") and hasattr(self.memory_encoder.out_proj, "weight"):
            self.mem_dim = self.memory_encoder.out_proj.weight.shape[0]
        self.num_maskmem = num_maskmem
        # keep the per-layer projected K/V of conditioning memories resident between frames (eval only);
        # MS2_MEMORY_BANK=0 selects the re-project-everything-per-frame path of the reference
        import os
        self.use_memory_bank_cache = os.environ.get("MS2_MEMORY_BANK", "1") != "0"
        # replay the fixed-shape per-slice sub-pipelines (SAM heads, memory encoder, batched image encoder) as
        # CUDA graphs (eval only; off by default, `++model.use_cuda_graphs=true` or MS2_CUDA_GRAPHS=1 turns it on)
        self.use_cuda_graphs = os.environ.get("MS2_CUDA_GRAPHS", "0") == "1"
        from ..runtime import GraphRunner
        self._graphs = GraphRunner()
        self._no_prompt = {}
        self.maskmem_tpos_enc = nn.Parameter(torch.zeros(num_maskmem, 1, 1, self.mem_dim))
        trunc_normal_(self.maskmem_tpos_enc, std=0.02)
        self.no_mem_embed = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        self.no_mem_pos_enc = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        trunc_normal_(self.no_mem_embed, std=0.02)
        trunc_normal_(self.no_mem_pos_enc, std=0.02)
        self.directly_add_no_mem_embed = directly_add_no_mem_embed
        self.sigmoid_scale_for_mem_enc = sigmoid_scale_for_mem_enc
        self.sigmoid_bias_for_mem_enc = sigmoid_bias_for_mem_enc
        self.binarize_mask_from_pts_for_mem_enc = binarize_mask_from_pts_for_mem_enc
        self.non_overlap_masks_for_mem_enc = non_overlap_masks_for_mem_enc
        self.memory_temporal_stride_for_eval = memory_temporal_stride_for_eval
        self.use_mask_input_as_output_without_sam = use_mask_input_as_output_without_sam
        self.multimask_output_in_sam = multimask_output_in_sam
        self.multimask_min_pt_num = multimask_min_pt_num
        self.multimask_max_pt_num = multimask_max_pt_num
        self.multimask_output_for_tracking = multimask_output_for_tracking
        self.use_multimask_token_for_obj_ptr = use_multimask_token_for_obj_ptr
        self.iou_prediction_use_sigmoid = iou_prediction_use_sigmoid
        self.image_size = image_size
        self.backbone_stride = backbone_stride
        self.sam_mask_decoder_extra_args = sam_mask_decoder_extra_args
        self.pred_obj_scores = pred_obj_scores
        self.pred_obj_scores_mlp = pred_obj_scores_mlp
        self.fixed_no_obj_ptr = fixed_no_obj_ptr
        self.soft_no_obj_ptr = soft_no_obj_ptr
        if self.fixed_no_obj_ptr:
            assert self.pred_obj_scores and self.use_obj_ptrs_in_encoder
        if self.pred_obj_scores and self.use_obj_ptrs_in_encoder:
            self.no_obj_ptr = nn.Parameter(torch.zeros(1, self.hidden_dim))
            trunc_normal_(self.no_obj_ptr, std=0.02)
        self.use_mlp_for_obj_ptr_proj = use_mlp_for_obj_ptr_proj
        self._build_sam_heads()
        self.add_all_frames_to_correct_as_cond = add_all_frames_to_correct_as_cond
        self.max_cond_frames_in_attn = max_cond_frames_in_attn
        assert not compile_image_encoder, "there is no tracing compiler on this path"

    @property
    def device(self):
        return next(self.parameters()).device

    # captured CUDA graphs bake in pointers to kernel-ready parameter copies: any event that can change the parameters
    # (checkpoint load, a train()/eval() round trip as in the reference's epoch loop, .to()/.cuda()) starts a new
    # parameter generation, and GraphRunner never replays a graph of an older one
    def load_state_dict(self, *args, **kwargs):
        from ..runtime import bump_generation
        bump_generation()
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode=True):
        from ..runtime import bump_generation
        if mode != self.training:
            bump_generation()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        from ..runtime import bump_generation
        bump_generation()
        return super()._apply(fn, *args, **kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Please use the corresponding methods in SAM2VideoPredictor for inference.")

    def _build_sam_heads(self):
        self.sam_prompt_embed_dim = self.hidden_dim
        self.sam_image_embedding_size = self.image_size // self.backbone_stride
        e = self.sam_image_embedding_size
        self.sam_prompt_encoder = PromptEncoder(embed_dim=self.sam_prompt_embed_dim, image_embedding_size=(e, e),
                                                input_image_size=(self.image_size, self.image_size), mask_in_chans=16)
        self.sam_mask_decoder = MaskDecoder(
            num_multimask_outputs=3,
            transformer=TwoWayTransformer(depth=2, embedding_dim=self.sam_prompt_embed_dim, mlp_dim=2048, num_heads=8),
            transformer_dim=self.sam_prompt_embed_dim, iou_head_depth=3, iou_head_hidden_dim=256,
            use_high_res_features=self.use_high_res_features_in_sam,
            iou_prediction_use_sigmoid=self.iou_prediction_use_sigmoid, pred_obj_scores=self.pred_obj_scores,
            pred_obj_scores_mlp=self.pred_obj_scores_mlp,
            use_multimask_token_for_obj_ptr=self.use_multimask_token_for_obj_ptr,
            **(self.sam_mask_decoder_extra_args or {}))
        if self.use_obj_ptrs_in_encoder:
            self.obj_ptr_proj = Linear(self.hidden_dim, self.hidden_dim)
            if self.use_mlp_for_obj_ptr_proj:
                self.obj_ptr_proj = MLP(self.hidden_dim, self.hidden_dim, self.hidden_dim, 3)
        else:
            self.obj_ptr_proj = nn.Identity()
        if self.proj_tpos_enc_in_obj_ptrs:
            self.obj_ptr_tpos_proj = Linear(self.hidden_dim, self.mem_dim)
        else:
            self.obj_ptr_tpos_proj = nn.Identity()

    # ------------------------------------------------------------------ image features
    def forward_image(self, img_batch):
        """sam2_base.py:464-476: encoder + conv_s0/conv_s1 on the two high-resolution levels."""
        def encode(img):
            d = self.sam_mask_decoder
            neck = self.image_encoder.neck
            # bf16 mode: conv_s0 / conv_s1 fold into the neck's lateral 1x1 convs of the two high-resolution levels
            # (no top-down sum there); the fp32 exact mode keeps the reference's two-step evaluation order
            if (self.use_high_res_features_in_sam and compute_dtype() == torch.bfloat16 and self.image_encoder.scalp >= 0
                    and 0 not in neck.fpn_top_down_levels and 1 not in neck.fpn_top_down_levels and len(neck.convs) >= 3):
                return self.image_encoder.forward_tokens(img, fold={0: d.conv_s0, 1: d.conv_s1})
            fs = self.image_encoder.forward_tokens(img)
            if self.use_high_res_features_in_sam:
                fs[0] = ops.gemm(to_compute(fs[0]), w_c(d.conv_s0.weight), p32(d.conv_s0.bias))
                fs[1] = ops.gemm(to_compute(fs[1]), w_c(d.conv_s1.weight), p32(d.conv_s1.bias))
            return fs
        # frames ingested as bf16 (utils/misc.py: streamed loaders in bf16 mode) feed the bf16 patch-embed im2col directly
        keep = img_batch.dtype == torch.bfloat16 and compute_dtype() == torch.bfloat16
        img = (img_batch if keep else img_batch.float()).contiguous()
        if self.use_cuda_graphs and not self.training and img.is_cuda and not torch.is_grad_enabled():
            feats = self._graphs.run(("image_encoder",), encode, [img])
        else:
            feats = encode(img)
        pe = self.image_encoder.neck.position_encoding
        B = img_batch.shape[0]
        fpn = [as_nchw_view(f) for f in feats]
        pos = [pe.table(f.shape[1], f.shape[2], f.device).permute(2, 0, 1)[None].expand(B, -1, -1, -1) for f in feats]
        return {"vision_features": fpn[-1], "vision_pos_enc": pos, "backbone_fpn": fpn}

    def _prepare_backbone_features(self, backbone_out):
        """sam2_base.py:478-492 (views only)."""
        backbone_out = backbone_out.copy()
        assert len(backbone_out["backbone_fpn"]) == len(backbone_out["vision_pos_enc"])
        assert len(backbone_out["backbone_fpn"]) >= self.num_feature_levels
        feature_maps = backbone_out["backbone_fpn"][-self.num_feature_levels:]
        vision_pos_embeds = backbone_out["vision_pos_enc"][-self.num_feature_levels:]
        feat_sizes = [(x.shape[-2], x.shape[-1]) for x in vision_pos_embeds]
        vision_feats = [x.flatten(2).permute(2, 0, 1) for x in feature_maps]
        vision_pos_embeds = [x.flatten(2).permute(2, 0, 1) for x in vision_pos_embeds]
        return backbone_out, vision_feats, vision_pos_embeds, feat_sizes

    # ------------------------------------------------------------------ SAM heads
    def _forward_sam_heads(self, backbone_features, point_inputs=None, mask_inputs=None, high_res_features=None,
                           multimask_output=False):
        """sam2_base.py:252-410; replayed as a CUDA graph per input signature when `use_cuda_graphs` is on."""
        if not (self.use_cuda_graphs and not self.training and backbone_features.is_cuda and
                not torch.is_grad_enabled()):
            return self._forward_sam_heads_impl(backbone_features, point_inputs, mask_inputs, high_res_features,
                                                multimask_output)
        pc = point_inputs["point_coords"] if point_inputs is not None else None
        pl = point_inputs["point_labels"] if point_inputs is not None else None
        hr = list(high_res_features) if high_res_features is not None else []

        def fn(bf, pc_, pl_, mi, *hr_):
            return self._forward_sam_heads_impl(bf, None if pc_ is None else {"point_coords": pc_, "point_labels": pl_},
                                                mi, list(hr_) if hr_ else None, multimask_output)
        return self._graphs.run(("sam_heads", bool(multimask_output), len(hr)), fn,
                                [backbone_features, pc, pl, mask_inputs] + hr)

    def _forward_sam_heads_impl(self, backbone_features, point_inputs=None, mask_inputs=None, high_res_features=None,
                                multimask_output=False):
        """sam2_base.py:252-410.  backbone_features NCHW-shaped [B,C,h,w]."""
        B = backbone_features.size(0)
        device = backbone_features.device
        assert backbone_features.size(1) == self.sam_prompt_embed_dim
        assert backbone_features.size(2) == self.sam_image_embedding_size
        assert backbone_features.size(3) == self.sam_image_embedding_size
        if point_inputs is not None:
            sam_point_coords = point_inputs["point_coords"]
            sam_point_labels = point_inputs["point_labels"]
            assert sam_point_coords.size(0) == B and sam_point_labels.size(0) == B
        else:
            # the "no prompt" padding point (sam2_base.py:300-303) is a constant: built once per (B, device)
            key = (B, str(device))
            if key not in self._no_prompt:
                self._no_prompt[key] = (torch.zeros(B, 1, 2, device=device),
                                        -torch.ones(B, 1, dtype=torch.int32, device=device))
            sam_point_coords, sam_point_labels = self._no_prompt[key]
        if mask_inputs is not None:
            assert len(mask_inputs.shape) == 4 and mask_inputs.shape[:2] == (B, 1)
            if tuple(mask_inputs.shape[-2:]) != tuple(self.sam_prompt_encoder.mask_input_size):
                sam_mask_prompt = ops.resize_bilinear(mask_inputs.float().contiguous(),
                                                      self.sam_prompt_encoder.mask_input_size, antialias=True)
            else:
                sam_mask_prompt = mask_inputs
        else:
            sam_mask_prompt = None
        # the decoder's output tokens ride in front of the sparse prompt rows (one buffer, no concatenation kernel)
        sparse, dense = self.sam_prompt_encoder(points=(sam_point_coords, sam_point_labels), boxes=None,
                                                masks=sam_mask_prompt, token_prefix=self.sam_mask_decoder._output_tokens())
        low_res_multimasks, ious, sam_output_tokens, object_score_logits = self.sam_mask_decoder(
            image_embeddings=backbone_features, image_pe=self.sam_prompt_encoder.get_dense_pe(),
            sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=multimask_output,
            repeat_image=False, high_res_features=high_res_features)
        low_res_multimasks = low_res_multimasks.float().contiguous()
        best = tok_best = None
        if multimask_output:
            # best mask by predicted IoU and (when the decoder returns one token per mask) its output token: one launch
            multi_tok = sam_output_tokens if sam_output_tokens.size(1) > 1 else None
            best, tok_best = ops.argmax_select_rows(ious.float(), None if multi_tok is None else multi_tok.float())
        obj_flat = object_score_logits.float().contiguous().view(-1)

        def mask_side():
            lr = low_res_multimasks
            if self.pred_obj_scores:
                lr = ops.gate_rows(lr, obj_flat, NO_OBJ_SCORE)
            hr = ops.resize_bilinear(lr, (self.image_size, self.image_size))
            if multimask_output:
                return lr, hr, ops.select_plane(lr, best), ops.select_plane(hr, best)
            return lr, hr, lr, hr

        def pointer_side():
            tok = tok_best if tok_best is not None else sam_output_tokens[:, 0]
            ptr = self.obj_ptr_proj(tok.contiguous())
            if self.pred_obj_scores:
                ptr = ops.obj_ptr_mix(ptr.float().contiguous(), obj_flat, p32(self.no_obj_ptr).view(-1),
                                      self.soft_no_obj_ptr, self.fixed_no_obj_ptr)
            return ptr

        # the mask path (gating, x4 up-sampling, plane selection) and the object pointer do not depend on each other
        (low_res_multimasks, high_res_multimasks, low_res_masks, high_res_masks), obj_ptr = par(mask_side, pointer_side)
        return (low_res_multimasks, high_res_multimasks, ious, low_res_masks, high_res_masks, obj_ptr,
                object_score_logits)

    def _use_mask_as_output(self, backbone_features, high_res_features, mask_inputs):
        """sam2_base.py:412-462."""
        out_scale, out_bias = 20.0, -10.0
        mask_f = mask_inputs.float().contiguous()
        B, _, H, W = mask_f.shape
        high_res_masks = ops.axpby(mask_f, out_scale, None, 0.0, out_bias)
        low_res_masks = ops.resize_bilinear(high_res_masks, (H // 4, W // 4), antialias=True)
        ious = mask_inputs.new_ones(B, 1).float()
        if not self.use_obj_ptrs_in_encoder:
            obj_ptr = torch.zeros(B, self.hidden_dim, device=mask_inputs.device)
        else:
            md = self.mask_downsample
            cols = ops.im2col(mask_f.view(B, H, W, 1), 4, 4, 0, compute_dtype())
            down = ops.gemm(cols, conv_w_c(md.weight), p32(md.bias), out_dtype=torch.float32)   # [B,H/4,W/4,1]
            _, _, _, _, _, obj_ptr, _ = self._forward_sam_heads(
                backbone_features=backbone_features, mask_inputs=down.view(B, 1, H // 4, W // 4),
                high_res_features=high_res_features)
        counts = ops.mask_stability_counts(mask_f, 0.0)
        # object present <=> any mask pixel > 0: logits = +10 / -10 (sam2_base.py:447-452)
        object_score_logits = ops.axpby(counts[:, 0:1].float().clamp_(max=1.0).contiguous(), out_scale, None, 0.0, out_bias)
        if self.pred_obj_scores:
            obj_ptr = ops.obj_ptr_mix(obj_ptr.float().contiguous(), object_score_logits.view(-1), p32(self.no_obj_ptr).view(-1),
                                      False, self.fixed_no_obj_ptr)
        return low_res_masks, high_res_masks, ious, low_res_masks, high_res_masks, obj_ptr, object_score_logits

    # ------------------------------------------------------------------ memory conditioning
    def _mem_pos_table(self, h, w, slot, device):
        """spatial sine-64 table + maskmem_tpos_enc[slot], token-major fp32 [h*w, mem_dim]
        (sam2_base.py:573-580); constant per parameter version."""
        pe = self.memory_encoder.position_encoding

        def make(tp):
            t = pe.table(h, w, device).reshape(h * w, -1)
            return (t + tp[slot].float().reshape(1, -1)).contiguous()
        return CACHE.get(self.maskmem_tpos_enc, ("mem_pos", h, w, slot, str(device)), make)

    def _memory_entries(self, frame_idx, output_dict, num_frames, track_in_reverse, B, device, lazy_ptrs=False):
        """sam2_base.py:518-637, un-concatenated: -> (cond, recent, ptrs, ptr_pos) where cond / recent are lists of
        (key, feats [B,hw,mem_dim] fp32 token-major, pos table [hw,mem_dim]) for the conditioning memories
        (t_pos = 0) and the <= num_maskmem-1 recent ones, and ptrs / ptr_pos are [B,n_tok,mem_dim] or None
        (lazy_ptrs: ptrs is a LIST of [B,C/mem_dim,mem_dim] views, one per pointer, ptr_pos None = zero)."""
        C = self.hidden_dim
        cond_outputs = output_dict["cond_frame_outputs"]
        assert len(cond_outputs) > 0
        selected, unselected = select_closest_cond_frames(frame_idx, cond_outputs, self.max_cond_frames_in_attn)
        t_pos_and_prevs = [(0, t, out) for t, out in selected.items()]
        r = self.memory_temporal_stride_for_eval
        for t_pos in range(1, self.num_maskmem):
            t_rel = self.num_maskmem - t_pos
            if t_rel == 1:
                prev_idx = frame_idx + t_rel if track_in_reverse else frame_idx - t_rel
            elif not track_in_reverse:
                prev_idx = ((frame_idx - 2) // r) * r - (t_rel - 2) * r
            else:
                prev_idx = -(-(frame_idx + 2) // r) * r + (t_rel - 2) * r
            out = output_dict["non_cond_frame_outputs"].get(prev_idx, None)
            if out is None:
                out = unselected.get(prev_idx, None)
            t_pos_and_prevs.append((t_pos, prev_idx, out))
        cond, recent = [], []
        for t_pos, t, prev in t_pos_and_prevs:
            if prev is None:
                continue
            src = prev["maskmem_features"]
            if src is None:
                # a conditioning memory that lives on another rank (split-KV bank, parallel.add_prompts_sharded): keep its
                # place in the list (ownership goes by position), MemoryAttention drops it before use
                assert t_pos == 0 and self.memory_attention.kv_shard is not None, "memory features are missing"
                cond.append(((t, None), None, None, None))
                continue
            # the per-memory item (token-major view + position table) depends on the stored frame output and the slot only:
            # cached ON that output (a tracked frame walks over every conditioning memory; 48 frames x 50 memories x a
            # handful of torch view calls was 12 % of the host time of a step)
            slot = self.num_maskmem - t_pos - 1
            ck = prev.get("_ms2_item")
            if ck is not None and ck[0] is src and ck[1] == (slot, B, str(device)):
                item = ck[2]
            else:
                feats = src if src.device == device else src.to(device, non_blocking=True)       # NCHW-shaped [B,Cm,h,w]
                h, w = feats.shape[-2:]
                item = ((t, id(src)), src, as_nhwc(feats.float()).reshape(B, h * w, self.mem_dim),
                        self._mem_pos_table(h, w, slot, device))
                prev["_ms2_item"] = (src, (slot, B, str(device)), item)
            (cond if t_pos == 0 else recent).append(item)
        ptrs = obj_pos = None
        if self.use_obj_ptrs_in_encoder:
            max_ptrs = min(num_frames, self.max_obj_ptrs_in_encoder)
            if not self.training and self.only_obj_ptrs_in_the_past_for_eval:
                ptr_cond = {t: o for t, o in selected.items() if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
            else:
                ptr_cond = selected
            ptr_outs = [(abs(frame_idx - t), o) for t, o in ptr_cond.items()]
            for t_diff in range(1, max_ptrs):
                t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
                if t < 0 or (num_frames is not None and t >= num_frames):
                    break
                o = output_dict["non_cond_frame_outputs"].get(t, unselected.get(t, None))
                if o is not None:
                    ptr_outs.append((t_diff, o))
            if ptr_outs and lazy_ptrs and not self.add_tpos_enc_to_obj_ptrs and self.mem_dim < C and C % self.mem_dim == 0:
                # banked path: the pointer tensors go into the bank's staging rows one by one (ms2_bank_rows), no
                # torch.stack / zeros / repeat_interleave; their position code is zero (sam2_base.py:627).  The
                # [B, C/mem_dim, mem_dim] view of a pointer is cached on its frame output.
                rr = C // self.mem_dim
                ptrs = []
                for _, o in ptr_outs:
                    v = o.get("_ms2_ptr_view")
                    if v is None or v[0] is not o["obj_ptr"]:
                        v = o["_ms2_ptr_view"] = (o["obj_ptr"], o["obj_ptr"].float().contiguous().view(B, rr, self.mem_dim))
                    ptrs.append(v[1])
                return cond, recent, ptrs, None
            pos_and_ptrs = [(d, o["obj_ptr"]) for d, o in ptr_outs]
            if pos_and_ptrs:
                pos_list, ptrs_list = zip(*pos_and_ptrs)
                ptrs = torch.stack([p.float() for p in ptrs_list], dim=1)                # [B,P,C]
                P = ptrs.shape[1]
                if self.add_tpos_enc_to_obj_ptrs:
                    from .sam2_utils import get_1d_sine_pe
                    t_diff_max = max_ptrs - 1
                    tpos_dim = C if self.proj_tpos_enc_in_obj_ptrs else self.mem_dim
                    obj_pos = get_1d_sine_pe(torch.tensor(pos_list, device=device) / t_diff_max, dim=tpos_dim)
                    if self.proj_tpos_enc_in_obj_ptrs:
                        obj_pos = self.obj_ptr_tpos_proj(obj_pos.contiguous())
                    obj_pos = obj_pos[None].expand(B, -1, -1)
                else:
                    obj_pos = ptrs.new_zeros(B, P, self.mem_dim)
                if self.mem_dim < C:
                    rr = C // self.mem_dim
                    ptrs = ptrs.reshape(B, P * rr, self.mem_dim)
                    obj_pos = obj_pos.repeat_interleave(rr, dim=1)
        return cond, recent, ptrs, obj_pos

    def _gather_memory(self, frame_idx, output_dict, num_frames, track_in_reverse, B, device):
        """sam2_base.py:518-637 -> (memory [B,Lk,mem_dim] fp32, memory_pos, num_obj_ptr_tokens)."""
        cond, recent, ptrs, obj_pos = self._memory_entries(frame_idx, output_dict, num_frames, track_in_reverse, B, device)
        mems = [e[2] for e in cond + recent]
        poss = [e[3][None].expand(B, -1, -1) for e in cond + recent]
        n_ptr_tok = 0
        if ptrs is not None:
            mems.append(ptrs)
            poss.append(obj_pos)
            n_ptr_tok = ptrs.shape[1]
        memory = torch.cat(mems, dim=1).contiguous()
        memory_pos = torch.cat(poss, dim=1).contiguous()
        return memory, memory_pos, n_ptr_tok

    def _prepare_memory_conditioned_features(self, frame_idx, is_init_cond_frame, current_vision_feats,
                                             current_vision_pos_embeds, feat_sizes, output_dict, num_frames,
                                             track_in_reverse=False):
        """sam2_base.py:494-663 -> NCHW-shaped [B,C,H,W]."""
        B = current_vision_feats[-1].size(1)
        C = self.hidden_dim
        H, W = feat_sizes[-1]
        device = current_vision_feats[-1].device
        curr = seq_to_tokens(current_vision_feats[-1].float())                     # [B,HW,C]
        if self.num_maskmem == 0:
            return as_nchw_view(curr.view(B, H, W, C))
        if is_init_cond_frame:
            if self.directly_add_no_mem_embed:
                out = ops.add_rowvec(curr, p32(self.no_mem_embed).view(-1))
                return as_nchw_view(out.view(B, H, W, C))
            memory = p32(self.no_mem_embed).expand(B, 1, self.mem_dim).contiguous()
            memory_pos = p32(self.no_mem_pos_enc).expand(B, 1, self.mem_dim).contiguous()
            n_ptr_tok = 0
        elif self.use_memory_bank_cache and not self.training:
            # conditioning memories are frame-invariant after projection: keep their per-layer
            # K (RoPE applied) / V resident and only project the <= 6 recent memories + object pointers per frame
            cond, recent, ptrs, obj_pos = self._memory_entries(frame_idx, output_dict, num_frames, track_in_reverse,
                                                               B, device, lazy_ptrs=True)
            bank = output_dict.get("_ms2_bank")
            if bank is None:
                from .memory_attention import MemoryBank
                bank = output_dict["_ms2_bank"] = MemoryBank()
            curr_pos = seq_to_tokens(current_vision_pos_embeds[-1].float())
            out = self.memory_attention.forward_tokens_banked(curr, curr_pos, bank, cond, recent, ptrs, obj_pos)
            return as_nchw_view(out.view(B, H, W, C))
        else:
            memory, memory_pos, n_ptr_tok = self._gather_memory(frame_idx, output_dict, num_frames, track_in_reverse,
                                                                B, device)
        curr_pos = seq_to_tokens(current_vision_pos_embeds[-1].float())
        out = self.memory_attention.forward_tokens(curr, curr_pos, memory, memory_pos, n_ptr_tok)
        return as_nchw_view(out.view(B, H, W, C))

    def _encode_new_memory(self, current_vision_feats, feat_sizes, pred_masks_high_res, is_mask_from_pts):
        """sam2_base.py:665-703; replayed as a CUDA graph per input signature when `use_cuda_graphs` is on."""
        if not (self.use_cuda_graphs and not self.training and pred_masks_high_res.is_cuda and
                not torch.is_grad_enabled()):
            return self._encode_new_memory_impl(current_vision_feats, feat_sizes, pred_masks_high_res, is_mask_from_pts)
        fs = [tuple(s) for s in feat_sizes]

        def fn(feat, masks):
            return self._encode_new_memory_impl([feat], fs, masks, is_mask_from_pts)
        return self._graphs.run(("mem_enc", bool(is_mask_from_pts), tuple(fs[-1])), fn,
                                [current_vision_feats[-1], pred_masks_high_res])

    def _encode_new_memory_impl(self, current_vision_feats, feat_sizes, pred_masks_high_res, is_mask_from_pts):
        """sam2_base.py:665-703 -> (maskmem_features NCHW-shaped, [pos NCHW-shaped])."""
        B = current_vision_feats[-1].size(1)
        C = self.hidden_dim
        H, W = feat_sizes[-1]
        pix = seq_to_tokens(current_vision_feats[-1].float()).view(B, H, W, C)
        if self.non_overlap_masks_for_mem_enc and not self.training:
            pred_masks_high_res = self._apply_non_overlapping_constraints(pred_masks_high_res)
        binarize = self.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts and not self.training
        m = pred_masks_high_res.float().contiguous()
        Hm, Wm = m.shape[-2:]
        feats = self.memory_encoder.forward_tokens(pix, m.view(B, Hm, Wm, 1), pre=2 if binarize else 1,
                                                   pre_scale=self.sigmoid_scale_for_mem_enc,
                                                   pre_bias=self.sigmoid_bias_for_mem_enc)
        feats_nchw = as_nchw_view(feats)
        pos = self.memory_encoder.position_encoding(feats_nchw)
        return feats_nchw, [pos]

    # ------------------------------------------------------------------ one tracking step
    def track_step(self, frame_idx, is_init_cond_frame, current_vision_feats, current_vision_pos_embeds, feat_sizes,
                   point_inputs, mask_inputs, output_dict, num_frames, track_in_reverse=False, run_mem_encoder=True,
                   prev_sam_mask_logits=None):
        """sam2_base.py:705-800."""
        current_out = {"point_inputs": point_inputs, "mask_inputs": mask_inputs}
        if len(current_vision_feats) > 1:
            high_res_features = [x.permute(1, 2, 0).view(x.size(1), x.size(2), *s)
                                 for x, s in zip(current_vision_feats[:-1], feat_sizes[:-1])]
        else:
            high_res_features = None
        if mask_inputs is not None and self.use_mask_input_as_output_without_sam:
            pix_feat = current_vision_feats[-1].permute(1, 2, 0)
            pix_feat = pix_feat.view(-1, self.hidden_dim, *feat_sizes[-1])
            sam_outputs = self._use_mask_as_output(pix_feat, high_res_features, mask_inputs)
        else:
            pix_feat_with_mem = self._prepare_memory_conditioned_features(
                frame_idx=frame_idx, is_init_cond_frame=is_init_cond_frame,
                current_vision_feats=current_vision_feats[-1:],
                current_vision_pos_embeds=current_vision_pos_embeds[-1:], feat_sizes=feat_sizes[-1:],
                output_dict=output_dict, num_frames=num_frames, track_in_reverse=track_in_reverse)
            if prev_sam_mask_logits is not None:
                assert point_inputs is not None and mask_inputs is None
                mask_inputs = prev_sam_mask_logits
            multimask_output = self._use_multimask(is_init_cond_frame, point_inputs)
            sam_outputs = self._forward_sam_heads(backbone_features=pix_feat_with_mem, point_inputs=point_inputs,
                                                  mask_inputs=mask_inputs, high_res_features=high_res_features,
                                                  multimask_output=multimask_output)
        _, _, _, low_res_masks, high_res_masks, obj_ptr, _ = sam_outputs
        current_out["pred_masks"] = low_res_masks
        current_out["pred_masks_high_res"] = high_res_masks
        current_out["obj_ptr"] = obj_ptr
        if run_mem_encoder and self.num_maskmem > 0:
            maskmem_features, maskmem_pos_enc = self._encode_new_memory(
                current_vision_feats=current_vision_feats, feat_sizes=feat_sizes,
                pred_masks_high_res=high_res_masks, is_mask_from_pts=(point_inputs is not None))
            current_out["maskmem_features"] = maskmem_features
            current_out["maskmem_pos_enc"] = maskmem_pos_enc
        else:
            current_out["maskmem_features"] = None
            current_out["maskmem_pos_enc"] = None
        return current_out

    def _use_multimask(self, is_init_cond_frame, point_inputs):
        """sam2_base.py:802-810."""
        num_pts = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        return (self.multimask_output_in_sam and (is_init_cond_frame or self.multimask_output_for_tracking)
                and (self.multimask_min_pt_num <= num_pts <= self.multimask_max_pt_num))

    def _apply_non_overlapping_constraints(self, pred_masks):
        """sam2_base.py:812-830: one kernel (ms2_non_overlap)."""
        if pred_masks.size(0) == 1:
            return pred_masks
        return ops.non_overlap(pred_masks.float().contiguous())
