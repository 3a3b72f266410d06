"""Pure-torch functional restatement of Medical-SAM2's per-slice inference hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA product path and
the `cpu_baseline` of bench.py — never the thing shipped.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the real reference from
/root/reference (with the 3-item runtime shim of SURVEY.md §8(c)), loads the same seeded
state_dict strictly, and dumps known-answer tensors into `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against them on every CPU run.

Everything is written against a flat `state_dict` (the reference checkpoint layout), so the
same weights feed the reference, this oracle and the product.  Each function cites the
reference file:line (relative to /root/reference/sam2_train/) it restates.
"""

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .config import get_config, hiera_blocks

NO_OBJ_SCORE = -1024.0  # modeling/sam2_base.py:19


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln2d(sd, name, x, eps=1e-6):
    """modeling/sam2_utils.py:133-147 (LayerNorm2d over the channel dim of NCHW)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[name + ".weight"][:, None, None] * x + sd[name + ".bias"][:, None, None]


def _mlp(sd, name, x, n, act=F.relu, sigmoid=False):
    """modeling/sam2_utils.py:107-130."""
    for i in range(n):
        x = _lin(sd, f"{name}.layers.{i}", x)
        if i < n - 1:
            x = act(x)
    return torch.sigmoid(x) if sigmoid else x


def sine_pos_enc(C, H, W, device, temperature=10000.0):
    """modeling/position_encoding.py:79-112 (normalize=True, scale=2*pi). Returns [C,H,W]."""
    npf = C // 2
    y = torch.arange(1, H + 1, dtype=torch.float32, device=device)
    x = torch.arange(1, W + 1, dtype=torch.float32, device=device)
    eps = 1e-6
    y = y / (y[-1] + eps) * (2 * math.pi)
    x = x / (x[-1] + eps) * (2 * math.pi)
    dim_t = torch.arange(npf, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    px = x[:, None] / dim_t            # [W, npf]
    py = y[:, None] / dim_t            # [H, npf]
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    pos = torch.cat((py[:, None, :].expand(H, W, npf), px[None, :, :].expand(H, W, npf)), dim=2)
    return pos.permute(2, 0, 1).contiguous()


def axial_rope_table(D, W, H, theta=10000.0, device="cpu"):
    """modeling/position_encoding.py:167-185. Returns (cos, sin) each [W*H, D/2]."""
    freqs = 1.0 / (theta ** (torch.arange(0, D, 4)[: D // 4].float() / D))
    t = torch.arange(W * H, dtype=torch.float32)
    tx = (t % W).float()
    ty = torch.div(t, W, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(tx, freqs), torch.outer(ty, freqs)], dim=-1)
    return ang.cos().to(device), ang.sin().to(device)


def apply_rope(x, cos, sin):
    """Real-valued restatement of apply_rotary_enc (position_encoding.py:188-216).
    x: [..., L, D]; cos/sin: [L, D/2]; adjacent channel pairs rotate."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xf[..., 0], xf[..., 1]
    out = torch.stack((a * cos - b * sin, a * sin + b * cos), dim=-1).flatten(-2)
    return out.type_as(x)


# --------------------------------------------------------------------------------------
# connected components (csrc/connected_components.cu) — label semantics of SURVEY §8(a) a12
# --------------------------------------------------------------------------------------
def connected_components_np(mask):
    """mask: uint8/bool ndarray [N,1,H,W] -> (labels int32, counts int32), same shape.

    label(p) = 1 + min over the 8-connected component of ((row&~1)*W + (col&~1)); 0 for
    background; counts = component area.  (connected_components.cu:62-209: block-based
    union-find on 2x2 blocks whose representative is the smallest block-anchor index.)
    """
    from scipy import ndimage
    mask = np.asarray(mask).astype(bool)
    N, _, H, W = mask.shape
    labels = np.zeros(mask.shape, np.int32)
    counts = np.zeros(mask.shape, np.int32)
    rows, cols = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    anchor = ((rows & ~1) * W + (cols & ~1)).astype(np.int64)
    for n in range(N):
        lab, num = ndimage.label(mask[n, 0], structure=np.ones((3, 3), np.int32))
        if num == 0:
            continue
        idx = np.arange(1, num + 1)
        mn = ndimage.minimum(anchor, lab, idx).astype(np.int64)
        area = ndimage.sum(np.ones_like(lab), lab, idx).astype(np.int64)
        lut_l = np.zeros(num + 1, np.int64)
        lut_c = np.zeros(num + 1, np.int64)
        lut_l[1:] = mn + 1
        lut_c[1:] = area
        labels[n, 0] = lut_l[lab]
        counts[n, 0] = lut_c[lab]
    return labels, counts


def fill_holes_in_mask_scores(mask, max_area):
    """utils/misc.py:247-258."""
    lab, area = connected_components_np((mask <= 0).cpu().numpy().astype(np.uint8))
    lab = torch.from_numpy(lab).to(mask.device)
    area = torch.from_numpy(area).to(mask.device)
    is_hole = (lab > 0) & (area <= max_area)
    return torch.where(is_hole, torch.full_like(mask, 0.1), mask)


def apply_non_overlapping_constraints(pred_masks):
    """modeling/sam2_base.py:812-830: per pixel only the object with the highest score keeps it, the others are
    clamped to <= -10 (ties: the FIRST object, torch.argmax)."""
    n = pred_masks.size(0)
    if n == 1:
        return pred_masks
    top = torch.argmax(pred_masks, dim=0, keepdim=True)
    keep = top == torch.arange(n, device=pred_masks.device)[:, None, None, None]
    return torch.where(keep, pred_masks, torch.clamp(pred_masks, max=-10.0))


def postprocess_masks(masks, orig_hw, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0):
    """utils/transforms.py:74-99.  Note the reference's order: the sprinkle test runs on the components of the
    ORIGINAL scores (`mask_flat` is taken before the holes are filled), its result is applied on top."""
    masks = masks.float()
    flat = masks.flatten(0, 1).unsqueeze(1)

    def cc(x):
        lab, area = connected_components_np(x.cpu().numpy().astype(np.uint8))
        return torch.from_numpy(lab).to(masks.device), torch.from_numpy(area).to(masks.device)
    if max_hole_area > 0:
        lab, area = cc(flat <= mask_threshold)
        hole = ((lab > 0) & (area <= max_hole_area)).reshape_as(masks)
        masks = torch.where(hole, torch.full_like(masks, mask_threshold + 10.0), masks)
    if max_sprinkle_area > 0:
        lab, area = cc(flat > mask_threshold)
        spr = ((lab > 0) & (area <= max_sprinkle_area)).reshape_as(masks)
        masks = torch.where(spr, torch.full_like(masks, mask_threshold - 10.0), masks)
    return F.interpolate(masks, tuple(orig_hw), mode="bilinear", align_corners=False)


def load_video_frames(video_path, image_size, img_mean=(0.485, 0.456, 0.406), img_std=(0.229, 0.224, 0.225)):
    """utils/misc.py:92-101,163-212: "<index>.jpg" files sorted by index; PIL decode -> RGB -> PIL resize (default
    resampling) -> /255 (float64, stored as fp32) -> (x - mean) / std.  -> (frames fp32 [T,3,S,S], video H, video W)."""
    import os
    from PIL import Image
    names = [p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG")]
    names.sort(key=lambda p: int(os.path.splitext(p)[0]))
    if not names:
        raise RuntimeError(f"no images found in {video_path}")
    frames = torch.zeros(len(names), 3, image_size, image_size, dtype=torch.float32)
    for n, name in enumerate(names):
        pil = Image.open(os.path.join(video_path, name))
        arr = np.array(pil.convert("RGB").resize((image_size, image_size))) / 255.0
        frames[n] = torch.from_numpy(arr).permute(2, 0, 1)
        w, h = pil.size
    frames -= torch.tensor(img_mean, dtype=torch.float32)[:, None, None]
    frames /= torch.tensor(img_std, dtype=torch.float32)[:, None, None]
    return frames, h, w


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class OracleSAM2:
    """Functional SAM2Base (modeling/sam2_base.py) over a flat state_dict."""

    def __init__(self, cfg, state_dict, device="cpu", dtype=torch.float32):
        if isinstance(cfg, str):
            cfg = get_config(cfg)
        self.cfg = cfg
        self.device = torch.device(device)
        self.sd = {k: v.to(self.device, dtype) for k, v in state_dict.items()}
        self.image_size = cfg["image_size"]
        self.hidden_dim = cfg["d_model"]
        self.mem_dim = cfg["mem_dim"]
        self.num_maskmem = cfg["num_maskmem"]
        self.emb_size = self.image_size // cfg["backbone_stride"]
        self.blocks, self.stage_ends = hiera_blocks(cfg)
        self._cache = {}

    # ---------------- image encoder ----------------
    def _hiera_pos_embed(self, h, w):
        """backbones/hieradet.py:269-277."""
        sd = self.sd
        pe = F.interpolate(sd["image_encoder.trunk.pos_embed"], size=(h, w), mode="bicubic")
        win = sd["image_encoder.trunk.pos_embed_window"]
        pe = pe + win.tile([x // y for x, y in zip(pe.shape, win.shape)])
        return pe.permute(0, 2, 3, 1)

    @staticmethod
    def _window_partition(x, ws):
        """backbones/utils.py:16-38."""
        B, H, W, C = x.shape
        ph = (ws - H % ws) % ws
        pw = (ws - W % ws) % ws
        if ph or pw:
            x = F.pad(x, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
        return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C), (Hp, Wp)

    @staticmethod
    def _window_unpartition(win, ws, pad_hw, hw):
        """backbones/utils.py:41-62."""
        Hp, Wp = pad_hw
        H, W = hw
        B = win.shape[0] // (Hp * Wp // ws // ws)
        x = win.view(B, Hp // ws, Wp // ws, ws, ws, -1)
        x = x.permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
        return x[:, :H, :W, :]

    @staticmethod
    def _pool(x):
        """backbones/hieradet.py:23-34 (MaxPool2d k2 s2 on NHWC)."""
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)

    def _ms_attention(self, p, b, x):
        """backbones/hieradet.py:58-83."""
        B, H, W, _ = x.shape
        nh = b["heads"]
        qkv = _lin(self.sd, p + "attn.qkv", x).reshape(B, H * W, 3, nh, -1)
        q, k, v = torch.unbind(qkv, 2)
        if b["q_pool"]:
            q = self._pool(q.reshape(B, H, W, -1))
            H, W = q.shape[1:3]
            q = q.reshape(B, H * W, nh, -1)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        o = o.transpose(1, 2).reshape(B, H, W, -1)
        return _lin(self.sd, p + "attn.proj", o)

    def _ms_block(self, i, x):
        """backbones/hieradet.py:136-168."""
        b = self.blocks[i]
        p = f"image_encoder.trunk.blocks.{i}."
        shortcut = x
        x = _ln(self.sd, p + "norm1", x, 1e-6)
        if b["dim"] != b["dim_out"]:
            shortcut = _lin(self.sd, p + "proj", x)
            if b["q_pool"]:
                shortcut = self._pool(shortcut)
        ws = b["window"]
        if ws > 0:
            H, W = x.shape[1], x.shape[2]
            x, pad_hw = self._window_partition(x, ws)
        x = self._ms_attention(p, b, x)
        if b["q_pool"]:
            ws = b["window"] // 2
            H, W = shortcut.shape[1:3]
            ph = (ws - H % ws) % ws
            pw = (ws - W % ws) % ws
            pad_hw = (H + ph, W + pw)
        if b["window"] > 0:
            x = self._window_unpartition(x, ws, pad_hw, (H, W))
        x = shortcut + x
        h = _ln(self.sd, p + "norm2", x, 1e-6)
        h = _lin(self.sd, p + "mlp.layers.1", F.gelu(_lin(self.sd, p + "mlp.layers.0", h)))
        return x + h

    def trunk(self, img):
        """backbones/hieradet.py:279-295 (+ PatchEmbed, backbones/utils.py:91-95)."""
        sd = self.sd
        x = F.conv2d(img, sd["image_encoder.trunk.patch_embed.proj.weight"],
                     sd["image_encoder.trunk.patch_embed.proj.bias"], stride=4, padding=3)
        x = x.permute(0, 2, 3, 1)
        x = x + self._hiera_pos_embed(x.shape[1], x.shape[2])
        outs = []
        for i in range(len(self.blocks)):
            x = self._ms_block(i, x)
            if i in self.stage_ends:
                outs.append(x.permute(0, 3, 1, 2))
        return outs

    def neck(self, xs):
        """backbones/image_encoder.py:101-133 (nearest top-down on levels 2,3; sum fuse)."""
        sd = self.sd
        n = len(xs) - 1
        out, pos = [None] * len(xs), [None] * len(xs)
        prev = None
        for i in range(n, -1, -1):
            w = sd[f"image_encoder.neck.convs.{n - i}.conv.weight"]
            b = sd[f"image_encoder.neck.convs.{n - i}.conv.bias"]
            lat = F.conv2d(xs[i], w, b)
            if i in self.cfg["fpn_top_down_levels"] and prev is not None:
                top = F.interpolate(prev.float(), scale_factor=2.0, mode="nearest")
                prev = lat + top
            else:
                prev = lat
            out[i] = prev
            pos[i] = sine_pos_enc(self.hidden_dim, prev.shape[-2], prev.shape[-1], prev.device)[None] \
                .expand(prev.shape[0], -1, -1, -1).to(prev.dtype)
        return out, pos

    def forward_image(self, img):
        """modeling/sam2_base.py:464-476 + backbones/image_encoder.py:29-42."""
        sd = self.sd
        feats, pos = self.neck(self.trunk(img))
        s = self.cfg["scalp"]
        feats, pos = feats[:-s], pos[:-s]
        feats = list(feats)
        feats[0] = F.conv2d(feats[0], sd["sam_mask_decoder.conv_s0.weight"], sd["sam_mask_decoder.conv_s0.bias"])
        feats[1] = F.conv2d(feats[1], sd["sam_mask_decoder.conv_s1.weight"], sd["sam_mask_decoder.conv_s1.bias"])
        return {"vision_features": feats[-1], "vision_pos_enc": list(pos), "backbone_fpn": feats}

    @staticmethod
    def prepare_backbone_features(backbone_out):
        """modeling/sam2_base.py:478-492 (num_feature_levels=3)."""
        fm = backbone_out["backbone_fpn"][-3:]
        pe = backbone_out["vision_pos_enc"][-3:]
        feat_sizes = [(x.shape[-2], x.shape[-1]) for x in pe]
        vf = [x.flatten(2).permute(2, 0, 1) for x in fm]
        vp = [x.flatten(2).permute(2, 0, 1) for x in pe]
        return backbone_out, vf, vp, feat_sizes

    # ---------------- memory attention ----------------
    def _rope_attn(self, p, q, k, v, num_k_exclude_rope=0, repeat_k=False):
        """modeling/sam/transformer.py:288-331 (1 head, D=256)."""
        sd = self.sd
        q = _lin(sd, p + ".q_proj", q)
        k = _lin(sd, p + ".k_proj", k)
        v = _lin(sd, p + ".v_proj", v)
        Lq = q.shape[-2]
        side = int(round(math.sqrt(Lq)))
        key = ("rope", q.shape[-1], side)
        if key not in self._cache:
            self._cache[key] = axial_rope_table(q.shape[-1], side, side, self.cfg["rope_theta"], q.device)
        cos, sin = self._cache[key]
        q = apply_rope(q, cos, sin)
        nk = k.shape[-2] - num_k_exclude_rope
        if nk > 0:
            r = nk // Lq if repeat_k else 1
            kr = apply_rope(k[..., :nk, :], cos.repeat(r, 1), sin.repeat(r, 1))
            k = torch.cat([kr, k[..., nk:, :]], dim=-2)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        return _lin(sd, p + ".out_proj", o)

    def memory_attention(self, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens=0):
        """modeling/memory_attention.py:58-169. Sequence-first in/out like the reference."""
        sd = self.sd
        if isinstance(curr, list):
            curr, curr_pos = curr[0], curr_pos[0]
        x = (curr + 0.1 * curr_pos).transpose(0, 1)
        mem = memory.transpose(0, 1)
        mem_pos = memory_pos.transpose(0, 1)
        for l in range(self.cfg["mem_attn_layers"]):
            p = f"memory_attention.layers.{l}."
            t = _ln(sd, p + "norm1", x, 1e-5)
            x = x + self._rope_attn(p + "self_attn", t, t, t)
            t = _ln(sd, p + "norm2", x, 1e-5)
            x = x + self._rope_attn(p + "cross_attn_image", t, mem + mem_pos, mem,
                                    num_k_exclude_rope=num_obj_ptr_tokens, repeat_k=True)
            t = _ln(sd, p + "norm3", x, 1e-5)
            x = x + _lin(sd, p + "linear2", F.relu(_lin(sd, p + "linear1", t)))
        x = _ln(sd, "memory_attention.norm", x, 1e-5)
        return x.transpose(0, 1)

    # ---------------- prompt encoder ----------------
    def _pe_encoding(self, coords01):
        """modeling/position_encoding.py:130-136."""
        g = self.sd["sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
        c = (2 * coords01 - 1) @ g
        c = 2 * np.pi * c
        return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)

    def get_dense_pe(self):
        """modeling/sam/prompt_encoder.py:68-77 + position_encoding.py:138-149."""
        h = w = self.emb_size
        dev = self.device
        grid = torch.ones((h, w), device=dev, dtype=torch.float32)
        y = (grid.cumsum(0) - 0.5) / h
        x = (grid.cumsum(1) - 0.5) / w
        pe = self._pe_encoding(torch.stack([x, y], dim=-1))
        return pe.permute(2, 0, 1).unsqueeze(0)

    def prompt_encoder(self, points, boxes, masks):
        """modeling/sam/prompt_encoder.py:79-190 with the (16,16) interpolate of :190 replaced
        by identity (SURVEY §8(c) shim iii). points=(coords[B,N,2], labels[B,N])."""
        sd = self.sd
        pfx = "sam_prompt_encoder."
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        sparse = torch.empty((bs, 0, self.hidden_dim), device=self.device)
        if points is not None:
            coords, labels = points
            coords = coords + 0.5
            if boxes is None:
                coords = torch.cat([coords, torch.zeros((bs, 1, 2), device=coords.device)], dim=1)
                labels = torch.cat([labels, -torch.ones((bs, 1), device=labels.device, dtype=labels.dtype)], dim=1)
            c = coords.clone().float()
            c[..., 0] = c[..., 0] / self.image_size
            c[..., 1] = c[..., 1] / self.image_size
            pe = self._pe_encoding(c)
            pe[labels == -1] = 0.0
            pe[labels == -1] += sd[pfx + "not_a_point_embed.weight"]
            for j in range(4):
                pe[labels == j] += sd[pfx + f"point_embeddings.{j}.weight"]
            sparse = torch.cat([sparse, pe], dim=1)
        if boxes is not None:
            c = (boxes + 0.5).reshape(-1, 2, 2) / self.image_size
            pe = self._pe_encoding(c.float())
            pe[:, 0, :] += sd[pfx + "point_embeddings.2.weight"]
            pe[:, 1, :] += sd[pfx + "point_embeddings.3.weight"]
            sparse = torch.cat([sparse, pe], dim=1)
        if masks is not None:
            x = F.conv2d(masks, sd[pfx + "mask_downscaling.0.weight"], sd[pfx + "mask_downscaling.0.bias"], stride=2)
            x = F.gelu(_ln2d(sd, pfx + "mask_downscaling.1", x))
            x = F.conv2d(x, sd[pfx + "mask_downscaling.3.weight"], sd[pfx + "mask_downscaling.3.bias"], stride=2)
            x = F.gelu(_ln2d(sd, pfx + "mask_downscaling.4", x))
            dense = F.conv2d(x, sd[pfx + "mask_downscaling.6.weight"], sd[pfx + "mask_downscaling.6.bias"])
        else:
            dense = sd[pfx + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bs, -1, self.emb_size, self.emb_size)
        return sparse, dense

    # ---------------- mask decoder ----------------
    def _attn(self, p, q, k, v, heads=8):
        """modeling/sam/transformer.py:199-263."""
        sd = self.sd
        q = _lin(sd, p + ".q_proj", q)
        k = _lin(sd, p + ".k_proj", k)
        v = _lin(sd, p + ".v_proj", v)

        def sep(x):
            b, n, c = x.shape
            return x.reshape(b, n, heads, c // heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sep(q), sep(k), sep(v))
        b, h, n, c = o.shape
        o = o.transpose(1, 2).reshape(b, n, h * c)
        return _lin(sd, p + ".out_proj", o)

    def _two_way_transformer(self, src, pos, tokens):
        """modeling/sam/transformer.py:74-196."""
        sd = self.sd
        keys = src.flatten(2).permute(0, 2, 1)
        key_pe = pos.flatten(2).permute(0, 2, 1)
        queries, query_pe = tokens, tokens
        for l in range(2):
            p = f"sam_mask_decoder.transformer.layers.{l}."
            if l == 0:
                queries = self._attn(p + "self_attn", queries, queries, queries)
            else:
                q = queries + query_pe
                queries = queries + self._attn(p + "self_attn", q, q, queries)
            queries = _ln(sd, p + "norm1", queries, 1e-5)
            q = queries + query_pe
            k = keys + key_pe
            queries = queries + self._attn(p + "cross_attn_token_to_image", q, k, keys)
            queries = _ln(sd, p + "norm2", queries, 1e-5)
            queries = queries + _lin(sd, p + "mlp.layers.1", F.relu(_lin(sd, p + "mlp.layers.0", queries)))
            queries = _ln(sd, p + "norm3", queries, 1e-5)
            q = queries + query_pe
            k = keys + key_pe
            keys = keys + self._attn(p + "cross_attn_image_to_token", k, q, queries)
            keys = _ln(sd, p + "norm4", keys, 1e-5)
        q = queries + query_pe
        k = keys + key_pe
        p = "sam_mask_decoder.transformer."
        queries = queries + self._attn(p + "final_attn_token_to_image", q, k, keys)
        queries = _ln(sd, p + "norm_final_attn", queries, 1e-5)
        return queries, keys

    def mask_decoder(self, image_embeddings, image_pe, sparse, dense, multimask_output,
                     repeat_image=False, cell_nums=None, high_res_features=None):
        """modeling/sam/mask_decoder.py:110-317 (cell_nums defaulting to None, shim ii)."""
        sd = self.sd
        pfx = "sam_mask_decoder."
        out_tok = torch.cat([sd[pfx + "obj_score_token.weight"], sd[pfx + "iou_token.weight"],
                             sd[pfx + "mask_tokens.weight"]], dim=0)
        out_tok = out_tok.unsqueeze(0).expand(sparse.size(0), -1, -1)
        tokens = torch.cat((out_tok, sparse), dim=1)
        if image_embeddings.size(0) != tokens.size(0) and cell_nums is not None:
            src = torch.repeat_interleave(image_embeddings, cell_nums, dim=0)
            pos_src = torch.repeat_interleave(image_pe, int(cell_nums.sum()), dim=0)
        else:
            src, pos_src = image_embeddings, image_pe
        src = src + dense
        b, c, h, w = src.shape
        hs, src = self._two_way_transformer(src, pos_src, tokens)
        iou_tok = hs[:, 1, :]
        mask_toks = hs[:, 2:6, :]
        src = src.transpose(1, 2).view(b, c, h, w)
        feat_s0, feat_s1 = high_res_features
        up = F.conv_transpose2d(src, sd[pfx + "output_upscaling.0.weight"], sd[pfx + "output_upscaling.0.bias"], stride=2)
        up = F.gelu(_ln2d(sd, pfx + "output_upscaling.1", up + feat_s1))
        up = F.conv_transpose2d(up, sd[pfx + "output_upscaling.3.weight"], sd[pfx + "output_upscaling.3.bias"], stride=2)
        up = F.gelu(up + feat_s0)
        hyper = torch.stack([_mlp(sd, pfx + f"output_hypernetworks_mlps.{i}", mask_toks[:, i, :], 3) for i in range(4)], dim=1)
        b, c, h, w = up.shape
        masks = (hyper @ up.view(b, c, h * w)).view(b, -1, h, w)
        iou_pred = _mlp(sd, pfx + "iou_prediction_head", iou_tok, 3, sigmoid=True)
        obj_logits = _mlp(sd, pfx + "pred_obj_score_head", hs[:, 0, :], 3)
        # forward(): mask_decoder.py:150-175
        if multimask_output:
            masks, iou_pred = masks[:, 1:], iou_pred[:, 1:]
        elif self.cfg["dynamic_multimask_via_stability"]:
            masks, iou_pred = self._dynamic_multimask(masks, iou_pred)
        else:
            masks, iou_pred = masks[:, 0:1], iou_pred[:, 0:1]
        sam_tokens = mask_toks[:, 1:] if multimask_output else mask_toks[:, 0:1]
        return masks, iou_pred, sam_tokens, obj_logits

    def _dynamic_multimask(self, all_logits, all_iou):
        """modeling/sam/mask_decoder.py:269-317."""
        d = self.cfg["dynamic_multimask_stability_delta"]
        mm, mi = all_logits[:, 1:], all_iou[:, 1:]
        best = torch.argmax(mi, dim=-1)
        bi = torch.arange(mi.size(0), device=mi.device)
        best_logits = mm[bi, best].unsqueeze(1)
        best_iou = mi[bi, best].unsqueeze(1)
        single, single_iou = all_logits[:, 0:1], all_iou[:, 0:1]
        flat = single.flatten(-2)
        ai = torch.sum(flat > d, dim=-1).float()
        au = torch.sum(flat > -d, dim=-1).float()
        stab = torch.where(au > 0, ai / au, torch.ones_like(au))
        ok = stab >= self.cfg["dynamic_multimask_stability_thresh"]
        return (torch.where(ok[..., None, None].expand_as(single), single, best_logits),
                torch.where(ok.expand_as(single_iou), single_iou, best_iou))

    # ---------------- SAM heads ----------------
    def forward_sam_heads(self, backbone_features, point_inputs=None, mask_inputs=None,
                          high_res_features=None, multimask_output=False):
        """modeling/sam2_base.py:252-410."""
        sd = self.sd
        B = backbone_features.size(0)
        dev = backbone_features.device
        if point_inputs is not None:
            pc, pl = point_inputs["point_coords"], point_inputs["point_labels"]
        else:
            pc = torch.zeros(B, 1, 2, device=dev)
            pl = -torch.ones(B, 1, dtype=torch.int32, device=dev)
        if mask_inputs is not None:
            ms = (4 * self.emb_size, 4 * self.emb_size)
            if tuple(mask_inputs.shape[-2:]) != ms:
                mask_prompt = F.interpolate(mask_inputs.float(), size=ms, align_corners=False,
                                            mode="bilinear", antialias=True)
            else:
                mask_prompt = mask_inputs
        else:
            mask_prompt = None
        sparse, dense = self.prompt_encoder((pc, pl), None, mask_prompt)
        low_mm, ious, sam_tokens, obj_logits = self.mask_decoder(
            backbone_features, self.get_dense_pe(), sparse, dense, multimask_output,
            high_res_features=high_res_features)
        is_obj = obj_logits > 0
        low_mm = torch.where(is_obj[:, None, None], low_mm, torch.full_like(low_mm, NO_OBJ_SCORE)).float()
        high_mm = F.interpolate(low_mm, size=(self.image_size, self.image_size), mode="bilinear", align_corners=False)
        tok = sam_tokens[:, 0]
        if multimask_output:
            best = torch.argmax(ious, dim=-1)
            bi = torch.arange(B, device=dev)
            low, high = low_mm[bi, best].unsqueeze(1), high_mm[bi, best].unsqueeze(1)
            if sam_tokens.size(1) > 1:
                tok = sam_tokens[bi, best]
        else:
            low, high = low_mm, high_mm
        obj_ptr = _mlp(sd, "obj_ptr_proj", tok, 3)
        lam = is_obj.float()
        obj_ptr = lam * obj_ptr + (1 - lam) * sd["no_obj_ptr"]
        return low_mm, high_mm, ious, low, high, obj_ptr, obj_logits

    def use_mask_as_output(self, backbone_features, high_res_features, mask_inputs):
        """modeling/sam2_base.py:412-462."""
        sd = self.sd
        mf = mask_inputs.float()
        high = mf * 20.0 - 10.0
        low = F.interpolate(high, size=(high.size(-2) // 4, high.size(-1) // 4), align_corners=False,
                            mode="bilinear", antialias=True)
        ious = mask_inputs.new_ones(mask_inputs.size(0), 1).float()
        md = F.conv2d(mf, sd["mask_downsample.weight"], sd["mask_downsample.bias"], stride=4)
        _, _, _, _, _, obj_ptr, _ = self.forward_sam_heads(backbone_features, mask_inputs=md,
                                                           high_res_features=high_res_features)
        is_obj = torch.any(mask_inputs.flatten(1).float() > 0.0, dim=1)[..., None]
        lam = is_obj.float()
        obj_logits = 20.0 * lam - 10.0
        obj_ptr = lam * obj_ptr + (1 - lam) * sd["no_obj_ptr"]
        return low, high, ious, low, high, obj_ptr, obj_logits

    # ---------------- memory encoder ----------------
    def memory_encoder(self, pix_feat, masks):
        """modeling/memory_encoder.py:158-181 with skip_mask_sigmoid=True."""
        sd = self.sd
        p = "memory_encoder."
        x = masks
        for j in range(4):
            x = F.conv2d(x, sd[p + f"mask_downsampler.encoder.{3 * j}.weight"],
                         sd[p + f"mask_downsampler.encoder.{3 * j}.bias"], stride=2, padding=1)
            x = F.gelu(_ln2d(sd, p + f"mask_downsampler.encoder.{3 * j + 1}", x))
        m = F.conv2d(x, sd[p + "mask_downsampler.encoder.12.weight"], sd[p + "mask_downsampler.encoder.12.bias"])
        x = F.conv2d(pix_feat, sd[p + "pix_feat_proj.weight"], sd[p + "pix_feat_proj.bias"]) + m
        for j in range(2):
            f = p + f"fuser.layers.{j}."
            inp = x
            x = F.conv2d(x, sd[f + "dwconv.weight"], sd[f + "dwconv.bias"], padding=3, groups=x.shape[1])
            x = _ln2d(sd, f + "norm", x).permute(0, 2, 3, 1)
            x = _lin(sd, f + "pwconv2", F.gelu(_lin(sd, f + "pwconv1", x)))
            x = (sd[f + "gamma"] * x).permute(0, 3, 1, 2)
            x = inp + x
        x = F.conv2d(x, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])
        pos = sine_pos_enc(self.mem_dim, x.shape[-2], x.shape[-1], x.device)[None].expand(x.shape[0], -1, -1, -1).to(x.dtype)
        return x, [pos]

    def encode_new_memory(self, current_vision_feats, feat_sizes, pred_masks_high_res, is_mask_from_pts):
        """modeling/sam2_base.py:665-703 (eval mode)."""
        B = current_vision_feats[-1].size(1)
        H, W = feat_sizes[-1]
        pix = current_vision_feats[-1].permute(1, 2, 0).reshape(B, self.hidden_dim, H, W)
        if self.cfg["binarize_mask_from_pts_for_mem_enc"] and is_mask_from_pts:
            m = (pred_masks_high_res > 0).float()
        else:
            m = torch.sigmoid(pred_masks_high_res)
        m = m * self.cfg["sigmoid_scale_for_mem_enc"] + self.cfg["sigmoid_bias_for_mem_enc"]
        return self.memory_encoder(pix, m)

    # ---------------- memory conditioning + track step ----------------
    def prepare_memory_conditioned_features(self, frame_idx, is_init_cond_frame, current_vision_feats,
                                            current_vision_pos_embeds, feat_sizes, output_dict, num_frames,
                                            track_in_reverse=False, return_memory=False):
        """modeling/sam2_base.py:494-663 (eval; max_cond_frames_in_attn=-1, stride 1,
        add_tpos_enc_to_obj_ptrs=False, only_obj_ptrs_in_the_past_for_eval=True)."""
        sd = self.sd
        B = current_vision_feats[-1].size(1)
        C = self.hidden_dim
        H, W = feat_sizes[-1]
        if is_init_cond_frame:
            x = current_vision_feats[-1] + sd["no_mem_embed"]
            return x.permute(1, 2, 0).reshape(B, C, H, W)
        to_cat, to_cat_pos = [], []
        cond = output_dict["cond_frame_outputs"]
        assert len(cond) > 0
        t_pos_and_prevs = [(0, out) for out in cond.values()]
        for t_pos in range(1, self.num_maskmem):
            t_rel = self.num_maskmem - t_pos
            prev_idx = frame_idx + t_rel if track_in_reverse else frame_idx - t_rel
            t_pos_and_prevs.append((t_pos, output_dict["non_cond_frame_outputs"].get(prev_idx, None)))
        for t_pos, prev in t_pos_and_prevs:
            if prev is None:
                continue
            feats = prev["maskmem_features"].to(self.device)
            to_cat.append(feats.flatten(2).permute(2, 0, 1))
            enc = prev["maskmem_pos_enc"][-1].to(self.device).flatten(2).permute(2, 0, 1)
            enc = enc + sd["maskmem_tpos_enc"][self.num_maskmem - t_pos - 1]
            to_cat_pos.append(enc)
        max_ptrs = min(num_frames, self.cfg["max_obj_ptrs_in_encoder"])
        ptr_cond = {t: o for t, o in cond.items() if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
        pos_and_ptrs = [(abs(frame_idx - t), o["obj_ptr"]) for t, o in ptr_cond.items()]
        for t_diff in range(1, max_ptrs):
            t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
            if t < 0 or (num_frames is not None and t >= num_frames):
                break
            o = output_dict["non_cond_frame_outputs"].get(t, None)
            if o is not None:
                pos_and_ptrs.append((t_diff, o["obj_ptr"]))
        n_ptr_tok = 0
        if pos_and_ptrs:
            ptrs = torch.stack([p for _, p in pos_and_ptrs], dim=0)            # [P,B,C]
            obj_pos = ptrs.new_zeros(len(pos_and_ptrs), B, self.mem_dim)
            r = C // self.mem_dim
            ptrs = ptrs.reshape(-1, B, r, self.mem_dim).permute(0, 2, 1, 3).flatten(0, 1)
            obj_pos = obj_pos.repeat_interleave(r, dim=0)
            to_cat.append(ptrs)
            to_cat_pos.append(obj_pos)
            n_ptr_tok = ptrs.shape[0]
        memory = torch.cat(to_cat, dim=0)
        memory_pos = torch.cat(to_cat_pos, dim=0)
        if return_memory:
            return memory, memory_pos, n_ptr_tok
        x = self.memory_attention(current_vision_feats, current_vision_pos_embeds, memory, memory_pos, n_ptr_tok)
        return x.permute(1, 2, 0).reshape(B, C, H, W)

    def use_multimask(self, is_init_cond_frame, point_inputs):
        """modeling/sam2_base.py:802-810 (multimask_output_in_sam and ..._for_tracking true)."""
        n = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        return self.cfg["multimask_min_pt_num"] <= n <= self.cfg["multimask_max_pt_num"]

    def track_step(self, frame_idx, is_init_cond_frame, current_vision_feats, current_vision_pos_embeds,
                   feat_sizes, point_inputs, mask_inputs, output_dict, num_frames, track_in_reverse=False,
                   run_mem_encoder=True, prev_sam_mask_logits=None):
        """modeling/sam2_base.py:705-800."""
        out = {"point_inputs": point_inputs, "mask_inputs": mask_inputs}
        hr = [x.permute(1, 2, 0).reshape(x.size(1), x.size(2), *s)
              for x, s in zip(current_vision_feats[:-1], feat_sizes[:-1])]
        if mask_inputs is not None:
            pix = current_vision_feats[-1].permute(1, 2, 0).reshape(-1, self.hidden_dim, *feat_sizes[-1])
            sam = self.use_mask_as_output(pix, hr, mask_inputs)
        else:
            pix = self.prepare_memory_conditioned_features(
                frame_idx, is_init_cond_frame, current_vision_feats[-1:], current_vision_pos_embeds[-1:],
                feat_sizes[-1:], output_dict, num_frames, track_in_reverse)
            if prev_sam_mask_logits is not None:
                mask_inputs = prev_sam_mask_logits
            sam = self.forward_sam_heads(pix, point_inputs, mask_inputs, hr,
                                         self.use_multimask(is_init_cond_frame, point_inputs))
        _, _, _, low, high, obj_ptr, _ = sam
        out["pred_masks"], out["pred_masks_high_res"], out["obj_ptr"] = low, high, obj_ptr
        if run_mem_encoder and self.num_maskmem > 0:
            f, pe = self.encode_new_memory(current_vision_feats, feat_sizes, high, point_inputs is not None)
            out["maskmem_features"], out["maskmem_pos_enc"] = f, pe
        else:
            out["maskmem_features"], out["maskmem_pos_enc"] = None, None
        return out


# --------------------------------------------------------------------------------------
# video predictor (sam2_video_predictor.py) — state machine restated for the flows that
# func_3d/function.py:226-274 drives: (train_)add_new_points/bbox/mask on fresh frames,
# then propagate_in_video forward.
# --------------------------------------------------------------------------------------
class OracleVideoPredictor:
    def __init__(self, model: OracleSAM2, fill_hole_area=None):
        self.m = model
        self.fill_hole_area = model.cfg["fill_hole_area"] if fill_hole_area is None else fill_hole_area

    def init_state(self, imgs_tensor, video_height=None, video_width=None):
        """sam2_video_predictor.py:107-176 + utils/misc.py:215-244."""
        m = self.m
        mean = torch.tensor((0.485, 0.456, 0.406), dtype=torch.float32)[:, None, None]
        std = torch.tensor((0.229, 0.224, 0.225), dtype=torch.float32)[:, None, None]
        images = (imgs_tensor.float() / 255.0).to(m.device)
        images = (images - mean.to(m.device)) / std.to(m.device)
        st = dict(images=images, num_frames=len(images),
                  video_height=video_height or m.image_size, video_width=video_width or m.image_size,
                  point_inputs_per_obj={}, mask_inputs_per_obj={}, cached_features={}, constants={},
                  obj_id_to_idx=OrderedDict(), obj_idx_to_id=OrderedDict(), obj_ids=[],
                  output_dict={"cond_frame_outputs": {}, "non_cond_frame_outputs": {}},
                  output_dict_per_obj={}, temp_output_dict_per_obj={},
                  consolidated_frame_inds={"cond_frame_outputs": set(), "non_cond_frame_outputs": set()},
                  tracking_has_started=False, frames_already_tracked={})
        self._image_feature(st, 0, 1)
        return st

    def _obj_idx(self, st, obj_id):
        """sam2_video_predictor.py:250-282."""
        idx = st["obj_id_to_idx"].get(obj_id)
        if idx is not None:
            return idx
        if st["tracking_has_started"]:
            raise RuntimeError(f"Cannot add new object id {obj_id} after tracking starts.")
        idx = len(st["obj_id_to_idx"])
        st["obj_id_to_idx"][obj_id] = idx
        st["obj_idx_to_id"][idx] = obj_id
        st["obj_ids"] = list(st["obj_id_to_idx"])
        st["point_inputs_per_obj"][idx] = {}
        st["mask_inputs_per_obj"][idx] = {}
        st["output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        st["temp_output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        return idx

    def reset_state(self, st):
        """sam2_video_predictor.py:1239-1268."""
        for k in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj",
                  "obj_id_to_idx", "obj_idx_to_id"):
            st[k].clear()
        st["obj_ids"] = []
        for d in (st["output_dict"], st["consolidated_frame_inds"]):
            d["cond_frame_outputs"].clear()
            d["non_cond_frame_outputs"].clear()
        st["tracking_has_started"] = False
        st["frames_already_tracked"].clear()

    def _image_feature(self, st, frame_idx, batch_size):
        """sam2_video_predictor.py:1270-1300 (1-frame cache)."""
        image, bo = st["cached_features"].get(frame_idx, (None, None))
        if bo is None:
            image = st["images"][frame_idx].float().unsqueeze(0)
            bo = self.m.forward_image(image)
            st["cached_features"] = {frame_idx: (image, bo)}
        ex = {"backbone_fpn": [f.expand(batch_size, -1, -1, -1) for f in bo["backbone_fpn"]],
              "vision_pos_enc": [p.expand(batch_size, -1, -1, -1) for p in bo["vision_pos_enc"]]}
        return self.m.prepare_backbone_features(ex)

    def _run_single_frame(self, st, output_dict, frame_idx, batch_size, is_init_cond_frame, point_inputs,
                          mask_inputs, reverse, run_mem_encoder, prev_sam_mask_logits=None):
        """sam2_video_predictor.py:1302-1367."""
        _, vf, vp, fs = self._image_feature(st, frame_idx, batch_size)
        out = self.m.track_step(frame_idx, is_init_cond_frame, vf, vp, fs, point_inputs, mask_inputs, output_dict,
                                st["num_frames"], reverse, run_mem_encoder, prev_sam_mask_logits)
        pred = out["pred_masks"]
        if self.fill_hole_area > 0:
            pred = fill_holes_in_mask_scores(pred, self.fill_hole_area)
        compact = {"maskmem_features": out["maskmem_features"],
                   "maskmem_pos_enc": self._maskmem_pos_enc(st, out),
                   "pred_masks": pred, "obj_ptr": out["obj_ptr"]}
        return compact, pred

    def _maskmem_pos_enc(self, st, out):
        """sam2_video_predictor.py:1399-1422."""
        pe = out["maskmem_pos_enc"]
        if pe is None:
            return None
        if "maskmem_pos_enc" not in st["constants"]:
            st["constants"]["maskmem_pos_enc"] = [x[0:1].clone() for x in pe]
        bs = pe[0].size(0)
        return [x.expand(bs, -1, -1, -1) for x in st["constants"]["maskmem_pos_enc"]]

    def add_new_points(self, st, frame_idx, obj_id, points, labels, clear_old_points=True, normalize_coords=True):
        """sam2_video_predictor.py:293-396."""
        m = self.m
        oi = self._obj_idx(st, obj_id)
        points = torch.as_tensor(points, dtype=torch.float32)
        labels = torch.as_tensor(labels, dtype=torch.int32)
        if points.dim() == 2:
            points = points.unsqueeze(0)
        if labels.dim() == 1:
            labels = labels.unsqueeze(0)
        if normalize_coords:
            points = points / torch.tensor([st["video_width"], st["video_height"]], dtype=torch.float32)
        points = (points * m.image_size).to(m.device)
        labels = labels.to(m.device)
        old = None if clear_old_points else st["point_inputs_per_obj"][oi].get(frame_idx)
        if old is not None:
            points = torch.cat([old["point_coords"], points], dim=1)
            labels = torch.cat([old["point_labels"], labels], dim=1)
        pin = {"point_coords": points, "point_labels": labels}
        st["point_inputs_per_obj"][oi][frame_idx] = pin
        st["mask_inputs_per_obj"][oi].pop(frame_idx, None)
        is_init = frame_idx not in st["frames_already_tracked"]
        reverse = False if is_init else st["frames_already_tracked"][frame_idx]["reverse"]
        od, tod = st["output_dict_per_obj"][oi], st["temp_output_dict_per_obj"][oi]
        key = "cond_frame_outputs" if is_init else "non_cond_frame_outputs"
        prev = tod[key].get(frame_idx) or od["cond_frame_outputs"].get(frame_idx) or od["non_cond_frame_outputs"].get(frame_idx)
        prev_logits = None
        if prev is not None and prev["pred_masks"] is not None:
            prev_logits = torch.clamp(prev["pred_masks"].to(m.device), -32.0, 32.0)
        out, _ = self._run_single_frame(st, od, frame_idx, 1, is_init, pin, None, reverse, False, prev_logits)
        tod[key][frame_idx] = out
        cons = self._consolidate(st, frame_idx, is_cond=is_init, run_mem_encoder=False, at_video_res=True)
        return frame_idx, st["obj_ids"], self._video_res(st, cons["pred_masks_video_res"])

    def add_new_bbox(self, st, frame_idx, obj_id, bbox, clear_old_points=True, normalize_coords=True):
        """sam2_video_predictor.py:399-422."""
        bbox = torch.as_tensor(bbox, dtype=torch.float32).reshape(-1, 2, 2)
        return self.add_new_points(st, frame_idx, obj_id, bbox, torch.tensor([2, 3], dtype=torch.int32),
                                   clear_old_points, normalize_coords)

    def add_new_mask(self, st, frame_idx, obj_id, mask):
        """sam2_video_predictor.py:557-638."""
        m = self.m
        oi = self._obj_idx(st, obj_id)
        mask = torch.as_tensor(mask)
        mo = mask[None, None].float().to(m.device)
        if mask.shape[0] != m.image_size or mask.shape[1] != m.image_size:
            mi = F.interpolate(mo, size=(m.image_size, m.image_size), align_corners=False, mode="bilinear", antialias=True)
            mi = (mi >= 0.5).float()
        else:
            mi = mo
        st["mask_inputs_per_obj"][oi][frame_idx] = mi
        st["point_inputs_per_obj"][oi].pop(frame_idx, None)
        is_init = frame_idx not in st["frames_already_tracked"]
        reverse = False if is_init else st["frames_already_tracked"][frame_idx]["reverse"]
        od, tod = st["output_dict_per_obj"][oi], st["temp_output_dict_per_obj"][oi]
        key = "cond_frame_outputs" if is_init else "non_cond_frame_outputs"
        out, _ = self._run_single_frame(st, od, frame_idx, 1, is_init, None, mi, reverse, False)
        tod[key][frame_idx] = out
        cons = self._consolidate(st, frame_idx, is_cond=is_init, run_mem_encoder=False, at_video_res=True)
        return frame_idx, st["obj_ids"], self._video_res(st, cons["pred_masks_video_res"])

    def _video_res(self, st, masks):
        """sam2_video_predictor.py:724-744."""
        hw = (st["video_height"], st["video_width"])
        if tuple(masks.shape[-2:]) == hw:
            return masks
        return F.interpolate(masks, size=hw, mode="bilinear", align_corners=False)

    def _consolidate(self, st, frame_idx, is_cond, run_mem_encoder, at_video_res=False):
        """sam2_video_predictor.py:746-862."""
        m = self.m
        bs = len(st["obj_idx_to_id"])
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if at_video_res:
            ch, cw, mk = st["video_height"], st["video_width"], "pred_masks_video_res"
        else:
            ch = cw = m.image_size // 4
            mk = "pred_masks"
        cons = {"maskmem_features": None, "maskmem_pos_enc": None,
                mk: torch.full((bs, 1, ch, cw), NO_OBJ_SCORE, dtype=torch.float32, device=m.device),
                "obj_ptr": torch.full((bs, m.hidden_dim), NO_OBJ_SCORE, dtype=torch.float32, device=m.device)}
        empty_ptr = None
        for oi in range(bs):
            out = st["temp_output_dict_per_obj"][oi][key].get(frame_idx)
            if out is None:
                out = st["output_dict_per_obj"][oi]["cond_frame_outputs"].get(frame_idx)
            if out is None:
                out = st["output_dict_per_obj"][oi]["non_cond_frame_outputs"].get(frame_idx)
            if out is None:
                if run_mem_encoder:
                    if empty_ptr is None:
                        empty_ptr = self._empty_mask_ptr(st, frame_idx)
                    cons["obj_ptr"][oi:oi + 1] = empty_ptr
                continue
            om = out["pred_masks"]
            if tuple(om.shape[-2:]) == (ch, cw):
                cons[mk][oi:oi + 1] = om
            else:
                cons[mk][oi:oi + 1] = F.interpolate(om, size=(ch, cw), mode="bilinear", align_corners=False)
            cons["obj_ptr"][oi:oi + 1] = out["obj_ptr"]
        if run_mem_encoder:
            high = F.interpolate(cons["pred_masks"], size=(m.image_size, m.image_size), mode="bilinear", align_corners=False)
            _, vf, _, fs = self._image_feature(st, frame_idx, bs)
            f, pe = m.encode_new_memory(vf, fs, high, True)
            cons["maskmem_features"] = f
            cons["maskmem_pos_enc"] = self._maskmem_pos_enc(st, {"maskmem_pos_enc": pe})
        return cons

    def _empty_mask_ptr(self, st, frame_idx):
        """sam2_video_predictor.py:864-898."""
        m = self.m
        mi = torch.zeros((1, 1, m.image_size, m.image_size), dtype=torch.float32, device=m.device)
        _, vf, vp, fs = self._image_feature(st, frame_idx, 1)
        out = m.track_step(frame_idx, True, vf, vp, fs, None, mi, {}, st["num_frames"], False, False, None)
        return out["obj_ptr"]

    def _add_output_per_object(self, st, frame_idx, out, key):
        """sam2_video_predictor.py:1210-1236."""
        for oi, od in st["output_dict_per_obj"].items():
            sl = slice(oi, oi + 1)
            o = {"maskmem_features": None, "maskmem_pos_enc": None,
                 "pred_masks": out["pred_masks"][sl], "obj_ptr": out["obj_ptr"][sl]}
            if out["maskmem_features"] is not None:
                o["maskmem_features"] = out["maskmem_features"][sl]
            if out["maskmem_pos_enc"] is not None:
                o["maskmem_pos_enc"] = [x[sl] for x in out["maskmem_pos_enc"]]
            od[key][frame_idx] = o

    def preflight(self, st):
        """sam2_video_predictor.py:901-968."""
        st["tracking_has_started"] = True
        od = st["output_dict"]
        cfi = st["consolidated_frame_inds"]
        for is_cond in (False, True):
            key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
            frames = set()
            for t in st["temp_output_dict_per_obj"].values():
                frames.update(t[key].keys())
            cfi[key].update(frames)
            for f in sorted(frames):
                cons = self._consolidate(st, f, is_cond=is_cond, run_mem_encoder=True)
                od[key][f] = cons
                self._add_output_per_object(st, f, cons, key)
            for t in st["temp_output_dict_per_obj"].values():
                t[key].clear()
        for f in od["cond_frame_outputs"]:
            od["non_cond_frame_outputs"].pop(f, None)
        for o in st["output_dict_per_obj"].values():
            for f in o["cond_frame_outputs"]:
                o["non_cond_frame_outputs"].pop(f, None)
        for f in cfi["cond_frame_outputs"]:
            cfi["non_cond_frame_outputs"].discard(f)

    def propagate_in_video(self, st, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """sam2_video_predictor.py:1041-1123 (generator)."""
        self.preflight(st)
        od = st["output_dict"]
        cfi = st["consolidated_frame_inds"]
        n = st["num_frames"]
        bs = len(st["obj_idx_to_id"])
        if not od["cond_frame_outputs"]:
            raise RuntimeError("No points are provided; please add points first")
        if start_frame_idx is None:
            start_frame_idx = min(od["cond_frame_outputs"])
        if max_frame_num_to_track is None:
            max_frame_num_to_track = n
        if reverse:
            end = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end - 1, -1) if start_frame_idx > 0 else []
        else:
            end = min(start_frame_idx + max_frame_num_to_track, n - 1)
            order = range(start_frame_idx, end + 1)
        for f in order:
            if f in cfi["cond_frame_outputs"]:
                key = "cond_frame_outputs"
                out = od[key][f]
                pred = out["pred_masks"]
            elif f in cfi["non_cond_frame_outputs"]:
                key = "non_cond_frame_outputs"
                out = od[key][f]
                pred = out["pred_masks"]
            else:
                key = "non_cond_frame_outputs"
                out, pred = self._run_single_frame(st, od, f, bs, False, None, None, reverse, True)
                od[key][f] = out
            self._add_output_per_object(st, f, out, key)
            st["frames_already_tracked"][f] = {"reverse": reverse}
            yield f, st["obj_ids"], self._video_res(st, pred)


# --------------------------------------------------------------------------------------
# image predictor (sam2_image_predictor.py)
# --------------------------------------------------------------------------------------
class OracleImagePredictor:
    def __init__(self, model: OracleSAM2, mask_threshold=0.0):
        self.m = model
        self.mask_threshold = mask_threshold
        self._features = None
        self._orig_hw = None

    def _transform(self, image):
        """utils/transforms.py:28-42 (ToTensor -> Resize(bilinear, antialias) -> Normalize)."""
        m = self.m
        x = torch.from_numpy(np.ascontiguousarray(image)).permute(2, 0, 1).float() / 255.0
        if x.shape[-2:] != (m.image_size, m.image_size):
            x = F.interpolate(x[None], size=(m.image_size, m.image_size), mode="bilinear",
                              align_corners=False, antialias=True)[0]
        mean = torch.tensor((0.485, 0.456, 0.406))[:, None, None]
        std = torch.tensor((0.229, 0.224, 0.225))[:, None, None]
        return (x - mean) / std

    def set_image_batch(self, image_list):
        """sam2_image_predictor.py:112-153."""
        m = self.m
        self._orig_hw = [im.shape[:2] for im in image_list]
        batch = torch.stack([self._transform(im) for im in image_list]).to(m.device)
        bo = m.forward_image(batch)
        _, vf, _, fs = m.prepare_backbone_features(bo)
        vf[-1] = vf[-1] + m.sd["no_mem_embed"]
        B = batch.shape[0]
        feats = [f.permute(1, 2, 0).reshape(B, -1, *s) for f, s in zip(vf, fs)]
        self._features = {"image_embed": feats[-1], "high_res_feats": feats[:-1]}

    def set_image(self, image):
        self.set_image_batch([image])

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None,
                multimask_output=True, return_logits=False, normalize_coords=True, img_idx=-1):
        """sam2_image_predictor.py:217-418."""
        m = self.m
        h, w = self._orig_hw[img_idx]
        concat = None
        if point_coords is not None:
            pc = torch.as_tensor(point_coords, dtype=torch.float, device=m.device).clone()
            if normalize_coords:
                pc[..., 0] = pc[..., 0] / w
                pc[..., 1] = pc[..., 1] / h
            pc = pc * m.image_size
            pl = torch.as_tensor(point_labels, dtype=torch.int, device=m.device)
            if pc.dim() == 2:
                pc, pl = pc[None], pl[None]
            concat = (pc, pl)
        if box is not None:
            bx = torch.as_tensor(box, dtype=torch.float, device=m.device).reshape(-1, 2, 2).clone()
            if normalize_coords:
                bx[..., 0] = bx[..., 0] / w
                bx[..., 1] = bx[..., 1] / h
            bx = bx * m.image_size
            bl = torch.tensor([[2, 3]], dtype=torch.int, device=m.device).repeat(bx.size(0), 1)
            concat = (torch.cat([bx, concat[0]], dim=1), torch.cat([bl, concat[1]], dim=1)) if concat else (bx, bl)
        if mask_input is not None:
            mask_input = torch.as_tensor(mask_input, dtype=torch.float, device=m.device)
            if mask_input.dim() == 3:
                mask_input = mask_input[None]
        sparse, dense = m.prompt_encoder(concat, None, mask_input)
        hr = [f[img_idx].unsqueeze(0) for f in self._features["high_res_feats"]]
        low, iou, _, _ = m.mask_decoder(self._features["image_embed"][img_idx].unsqueeze(0), m.get_dense_pe(),
                                        sparse, dense, multimask_output, high_res_features=hr)
        masks = F.interpolate(low.float(), (h, w), mode="bilinear", align_corners=False)
        low = torch.clamp(low, -32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return (masks.squeeze(0).float().cpu().numpy(), iou.squeeze(0).float().cpu().numpy(),
                low.squeeze(0).float().cpu().numpy())

    def predict_batch(self, point_coords_batch=None, point_labels_batch=None, box_batch=None,
                      mask_input_batch=None, multimask_output=True, return_logits=False, normalize_coords=True):
        """sam2_image_predictor.py:155-215."""
        outs = ([], [], [])
        for i in range(len(self._orig_hw)):
            r = self.predict(None if point_coords_batch is None else point_coords_batch[i],
                             None if point_labels_batch is None else point_labels_batch[i],
                             None if box_batch is None else box_batch[i],
                             None if mask_input_batch is None else mask_input_batch[i],
                             multimask_output, return_logits, normalize_coords, img_idx=i)
            for o, x in zip(outs, r):
                o.append(x)
        return outs
