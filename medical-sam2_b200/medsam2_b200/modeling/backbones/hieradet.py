"""Hiera trunk on token-major (NHWC) fp32 streams with compute-dtype GEMM/attention operands.

Same parameters / state_dict keys as the reference trunk (backbones/hieradet.py:171-260), but the
data flow is re-designed: no window_partition / F.pad / window_unpartition copies and no padded qkv
GEMM — the qkv Linear runs on the un-padded tokens and `ms2_window_attention` gathers windows,
substitutes the qkv bias for zero-padded positions (they are real keys in the reference),
max-pools q and scatters the cropped output; residual adds live in the GEMM epilogues; the
bicubic+tiled position embedding is a cached table added inside the patch-embed kernel.
"""
import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ...runtime import CACHE, compute_dtype, p32, w_c
from ..sam2_utils import LayerNorm, Linear, MLP, as_nchw_view


class PatchEmbed(nn.Module):
    def __init__(self, kernel_size=(7, 7), stride=(4, 4), padding=(3, 3), in_chans=3, embed_dim=768):
        super().__init__()
        assert tuple(kernel_size) == (7, 7) and tuple(stride) == (4, 4) and tuple(padding) == (3, 3) and in_chans == 3
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x, pos=None):
        """x fp32 NCHW -> NHWC tokens (+pos table).  bf16 mode: im2col -> tensor-core GEMM (bias and the pos-embed
        table ride in the epilogue); fp32 mode: the direct fp32 conv kernel."""
        if compute_dtype() != torch.bfloat16:
            return ops.patch_embed(x, p32(self.proj.weight), p32(self.proj.bias), pos)
        B, _, Hin, Win = x.shape
        Ho, Wo = (Hin + 6 - 7) // 4 + 1, (Win + 6 - 7) // 4 + 1
        E = self.proj.weight.shape[0]
        w = CACHE.get(self.proj.weight, "patch_w152",
                      lambda t: torch.nn.functional.pad(t.detach().permute(0, 2, 3, 1).reshape(t.shape[0], -1), (0, 5))
                      .to(torch.bfloat16).contiguous())
        cols = ops.patch_im2col(x)
        out = torch.empty((B, Ho, Wo, E), dtype=torch.float32, device=x.device)
        res = None if pos is None else pos.reshape(Ho * Wo, E)
        for b in range(B):
            ops.gemm(cols[b], w, p32(self.proj.bias), residual=res, out=out[b].view(Ho * Wo, E))
        return out


class MultiScaleAttention(nn.Module):
    def __init__(self, dim, dim_out, num_heads, q_pool=False):
        super().__init__()
        self.dim, self.dim_out, self.num_heads = dim, dim_out, num_heads
        self.q_pool = q_pool
        self.qkv = Linear(dim, dim_out * 3)
        self.proj = Linear(dim_out, dim_out)

    def forward(self, xn, window, shortcut):
        """xn: LayerNormed tokens [B,H,W,dim] (compute dtype); returns shortcut + proj(attn) fp32."""
        B, H, W, _ = xn.shape
        nh, D = self.num_heads, self.dim_out // self.num_heads
        qkv = self.qkv(xn, out_dtype=compute_dtype())
        if window > 0 or self.q_pool:
            ws = window if window > 0 else 2 * ((max(H, W) + 1) // 2)
            o = ops.window_attention(qkv, p32(self.qkv.bias), B, H, W, nh, D, ws, self.q_pool)
        else:
            t = qkv.view(B, H * W, 3 * self.dim_out)
            o = ops.attention(t[:, :, : self.dim_out], t[:, :, self.dim_out: 2 * self.dim_out],
                              t[:, :, 2 * self.dim_out:], nh).view(B, H, W, self.dim_out)
        return self.proj(o, out_dtype=torch.float32, residual=shortcut)


class MultiScaleBlock(nn.Module):
    def __init__(self, dim, dim_out, num_heads, mlp_ratio=4.0, drop_path=0.0, norm_layer="LayerNorm",
                 q_stride=None, act_layer=None, window_size=0):
        super().__init__()
        self.dim, self.dim_out = dim, dim_out
        self.norm1 = LayerNorm(dim, eps=1e-6)
        self.window_size = window_size
        self.q_stride = q_stride
        if q_stride:
            assert tuple(q_stride) == (2, 2)
        self.attn = MultiScaleAttention(dim, dim_out, num_heads=num_heads, q_pool=bool(q_stride))
        self.norm2 = LayerNorm(dim_out, eps=1e-6)
        self.mlp = MLP(dim_out, int(dim_out * mlp_ratio), dim_out, num_layers=2, activation="gelu")
        if dim != dim_out:
            self.proj = Linear(dim, dim_out)

    def forward(self, x):
        """x fp32 [B,H,W,dim] -> fp32 [B,H',W',dim_out]   (hieradet.py:136-168)."""
        cd = compute_dtype()
        xn = self.norm1(x, out_dtype=cd)
        shortcut = x
        if self.dim != self.dim_out:
            shortcut = self.proj(xn, out_dtype=torch.float32)
            if self.q_stride:
                shortcut = ops.maxpool2x2(shortcut)
        x = self.attn(xn, self.window_size, shortcut)
        h = self.norm2(x, out_dtype=cd)
        return self.mlp(h, residual=x)


class Hiera(nn.Module):
    def __init__(self, embed_dim=96, num_heads=1, drop_path_rate=0.0, q_pool=3, q_stride=(2, 2),
                 stages=(2, 3, 16, 3), dim_mul=2.0, head_mul=2.0, window_pos_embed_bkg_spatial_size=(14, 14),
                 window_spec=(8, 4, 14, 7), global_att_blocks=(12, 16, 20), return_interm_layers=True):
        super().__init__()
        assert len(stages) == len(window_spec)
        self.window_spec = tuple(window_spec)
        depth = sum(stages)
        self.q_stride = tuple(q_stride)
        self.stage_ends = [sum(stages[:i]) - 1 for i in range(1, len(stages) + 1)]
        assert 0 <= q_pool <= len(self.stage_ends[:-1])
        self.q_pool_blocks = [x + 1 for x in self.stage_ends[:-1]][:q_pool]
        self.return_interm_layers = return_interm_layers
        self.patch_embed = PatchEmbed(embed_dim=embed_dim)
        self.global_att_blocks = tuple(global_att_blocks or ())
        self.window_pos_embed_bkg_spatial_size = tuple(window_pos_embed_bkg_spatial_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim, *self.window_pos_embed_bkg_spatial_size))
        self.pos_embed_window = nn.Parameter(torch.zeros(1, embed_dim, self.window_spec[0], self.window_spec[0]))
        cur_stage = 1
        self.blocks = nn.ModuleList()
        for i in range(depth):
            dim_out = embed_dim
            window_size = self.window_spec[cur_stage - 1]   # lags one block at stage changes
            if i in self.global_att_blocks:
                window_size = 0
            if i - 1 in self.stage_ends:
                dim_out = int(embed_dim * dim_mul)
                num_heads = int(num_heads * head_mul)
                cur_stage += 1
            self.blocks.append(MultiScaleBlock(dim=embed_dim, dim_out=dim_out, num_heads=num_heads,
                                               q_stride=self.q_stride if i in self.q_pool_blocks else None,
                                               window_size=window_size))
            embed_dim = dim_out
        self.channel_list = ([self.blocks[i].dim_out for i in self.stage_ends[::-1]]
                             if return_interm_layers else [self.blocks[-1].dim_out])

    def _pos_table(self, h, w):
        """bicubic(pos_embed) + tiled window embed as an fp32 [h,w,C] table (hieradet.py:269-277);
        input-independent, so computed once per parameter version instead of every forward."""
        def make(pe, win):
            t = F.interpolate(pe.float(), size=(h, w), mode="bicubic")
            t = t + win.float().tile([x // y for x, y in zip(t.shape, win.shape)])
            return t[0].permute(1, 2, 0).contiguous()
        return CACHE.get((self.pos_embed, self.pos_embed_window), ("hiera_pos", h, w), make)

    def forward_tokens(self, img):
        """img fp32 NCHW -> list of NHWC fp32 stage outputs."""
        Hp, Wp = (img.shape[-2] + 6 - 7) // 4 + 1, (img.shape[-1] + 6 - 7) // 4 + 1
        x = self.patch_embed(img, self._pos_table(Hp, Wp))
        outs = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i == self.stage_ends[-1] or (i in self.stage_ends and self.return_interm_layers):
                outs.append(x)
        return outs

    def forward(self, x):
        return [as_nchw_view(t) for t in self.forward_tokens(x)]
