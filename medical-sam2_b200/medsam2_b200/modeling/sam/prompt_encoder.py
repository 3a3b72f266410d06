"""SAM prompt encoder (reference modeling/sam/prompt_encoder.py:17-190) on the native kernels.

Sparse prompts: random-Fourier features of the click / box-corner coordinates (ms2_fourier_pe) plus
a label-indexed embedding row; dense prompts: the two 2x2/s2 convs as im2col+GEMM with the
LayerNorm2d+GELU fused, or the `no_mask_embed` row broadcast.  The dense output keeps the spatial
size of the embedding grid (the fork's hard-coded (16,16) interpolate at :190 is the identity at the
configured image size and is not reproduced; SURVEY.md §0 finding 1).
"""
import torch
from torch import nn

from ... import ops
from ...runtime import CACHE, compute_dtype, conv_w_c, p32, w_c
from ..sam2_utils import LayerNorm2d, as_nchw_view, as_nhwc


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=64, scale=None):
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def encode(self, coords01):
        """coords normalised to [0,1], [...,2] -> [...,2F] = [sin | cos]."""
        return ops.fourier_pe(coords01.float().contiguous(), p32(self.positional_encoding_gaussian_matrix))

    def grid_table(self, h, w):
        """token-major [h,w,2F] encoding of the pixel-centre grid (position_encoding.py:138-149); cached."""
        g = self.positional_encoding_gaussian_matrix

        def make(gm):
            ys = (torch.arange(h, device=gm.device, dtype=torch.float32) + 0.5) / h
            xs = (torch.arange(w, device=gm.device, dtype=torch.float32) + 0.5) / w
            grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1).contiguous()
            return ops.fourier_pe(grid, gm.float().contiguous())
        return CACHE.get(g, ("pe_grid", h, w), make)

    def forward(self, size):
        h, w = size
        return self.grid_table(h, w).permute(2, 0, 1)

    def forward_with_coords(self, coords_input, image_size):
        c = coords_input.clone().float()
        c[:, :, 0] = c[:, :, 0] / image_size[1]
        c[:, :, 1] = c[:, :, 1] / image_size[0]
        return self.encode(c)


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = tuple(input_image_size)
        self.image_embedding_size = tuple(image_embedding_size)
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4
        self.point_embeddings = nn.ModuleList(nn.Embedding(1, embed_dim) for _ in range(4))
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans // 4),
            nn.Identity(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans),
            nn.Identity(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1),
        )
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self):
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def dense_pe_tokens(self):
        h, w = self.image_embedding_size
        return self.pe_layer.grid_table(h, w).view(1, h * w, self.embed_dim)

    def _label_table(self):
        ps = (self.not_a_point_embed.weight,) + tuple(e.weight for e in self.point_embeddings)
        return CACHE.get(ps, "label_table", lambda *ts: torch.cat([t.float() for t in ts], 0).contiguous())

    def _embed_points(self, points, labels, pad, prefix=None):
        if points.dim() == 3 and labels.dim() == 2 and not labels.dtype.is_floating_point:
            return ops.point_embed(points, labels, p32(self.pe_layer.positional_encoding_gaussian_matrix),
                                   self._label_table(), pad, self.input_image_size, prefix)
        assert prefix is None
        points = points.float() + 0.5
        if pad:
            points = torch.cat([points, torch.zeros((points.shape[0], 1, 2), device=points.device)], dim=1)
            labels = torch.cat([labels, -torch.ones((labels.shape[0], 1), device=labels.device, dtype=labels.dtype)], dim=1)
        pe = self.pe_layer.forward_with_coords(points, self.input_image_size)
        lab = labels.long()
        keep = (lab != -1).to(pe.dtype).unsqueeze(-1)
        return pe * keep + self._label_table()[(lab + 1).clamp(0, 4)]

    def _embed_boxes(self, boxes):
        coords = (boxes.float() + 0.5).reshape(-1, 2, 2)
        pe = self.pe_layer.forward_with_coords(coords, self.input_image_size)
        t = self._label_table()
        return pe + torch.stack([t[3], t[4]], 0)[None]

    def embed_masks_tokens(self, masks_nhwc):
        cd = compute_dtype()
        c0, n0, _, c1, n1, _, c2 = self.mask_downscaling
        x = ops.gemm(ops.im2col(masks_nhwc, 2, 2, 0, cd), conv_w_c(c0.weight), p32(c0.bias), out_dtype=torch.float32)
        x = n0(x, act=ops.ACT_GELU)
        x = ops.gemm(ops.im2col(x, 2, 2, 0, cd), conv_w_c(c1.weight), p32(c1.bias), out_dtype=torch.float32)
        x = n1(x, out_dtype=cd, act=ops.ACT_GELU)
        return ops.gemm(x, w_c(c2.weight), p32(c2.bias), out_dtype=torch.float32)

    def forward(self, points, boxes, masks, batch_size=-1, token_prefix=None):
        """-> (sparse [B,N,C] fp32, dense NCHW-shaped [B,C,h,w] fp32).  `token_prefix` (fp32 [P,C], optional, not in the
        reference): the click embeddings are written behind P constant rows of ONE buffer and `sparse` is the view past
        them; the buffer rides along as `sparse._ms2_tokens` so that the mask decoder can use it as its token matrix
        (output tokens + sparse prompts, mask_decoder.py:186-196) without a concatenation kernel."""
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        dev = self.no_mask_embed.weight.device
        sparse = None
        if points is not None:
            coords, labels = points
            fused = (token_prefix is not None and boxes is None and coords.dim() == 3 and labels.dim() == 2
                     and not labels.dtype.is_floating_point)
            emb = self._embed_points(coords, labels, pad=(boxes is None), prefix=token_prefix if fused else None)
            if fused:
                sparse = emb[:, token_prefix.shape[0]:]
                sparse._ms2_tokens = emb
            else:
                sparse = emb
        if boxes is not None:
            be = self._embed_boxes(boxes)
            sparse = be if sparse is None else torch.cat([sparse, be], dim=1)
        if sparse is None:
            sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        if masks is not None:
            B, _, H, W = masks.shape
            dense = as_nchw_view(self.embed_masks_tokens(masks.float().contiguous().view(B, H, W, 1)))
        else:
            dense = p32(self.no_mask_embed.weight).reshape(1, -1, 1, 1).expand(
                bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense
