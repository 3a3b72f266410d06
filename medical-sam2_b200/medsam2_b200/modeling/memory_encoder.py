"""Memory encoder (reference modeling/memory_encoder.py:17-181), token-major.

mask -> 4x {3x3/s2 conv as im2col+GEMM, LayerNorm2d+GELU fused in the norm kernel} -> 1x1;
pix_feat_proj GEMM with the mask features added as the GEMM residual; 2x CXBlock = depth-wise 7x7
kernel -> LN -> GEMM+GELU -> GEMM with layer-scale and the block residual in the epilogue;
out_proj GEMM to 64 channels.  The sigmoid / binarise + scale/bias of sam2_base.py:686-696 is fused
into the first im2col gather (zero padding applies AFTER that transform, as in the reference).
"""
import copy

import torch
from torch import nn

from .. import ops
from ..runtime import compute_dtype, conv_w_c, p32, w_c
from .sam2_utils import LayerNorm2d, Linear, as_nchw_view, as_nhwc, to_compute


class MaskDownSampler(nn.Module):
    def __init__(self, embed_dim=256, kernel_size=4, stride=4, padding=0, total_stride=16, activation=None):
        super().__init__()
        import math
        num_layers = int(math.log2(total_stride) // math.log2(stride))
        assert stride ** num_layers == total_stride
        self.kernel_size, self.stride, self.padding, self.num_layers = kernel_size, stride, padding, num_layers
        self.encoder = nn.Sequential()
        cin = 1
        for _ in range(num_layers):
            cout = cin * (stride ** 2)
            self.encoder.append(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=padding))
            self.encoder.append(LayerNorm2d(cout))
            self.encoder.append(nn.Identity())   # GELU slot (fused into the norm kernel); keeps indices 0..11
            cin = cout
        self.encoder.append(nn.Conv2d(cin, embed_dim, kernel_size=1))

    def forward_tokens(self, x_nhwc, pre=0, pre_scale=1.0, pre_bias=0.0):
        cd = compute_dtype()
        x = x_nhwc
        for j in range(self.num_layers):
            conv, norm = self.encoder[3 * j], self.encoder[3 * j + 1]
            last = j == self.num_layers - 1
            if (self.kernel_size, self.stride, self.padding) == (3, 2, 1) and \
                    (conv.in_channels, conv.out_channels) in _THIN_CONVS:
                # thin layers: one fused direct-conv + LN2d + GELU kernel instead of im2col + GEMM + LN (the 16 -> 64 layer,
                # 151 MFLOP at 128x128, takes 68 us as a direct SIMT conv and ~20 us as im2col + tcgen05 GEMM + LN)
                x = ops.conv3x3s2_ln_gelu(x, p32(conv.weight), p32(conv.bias), p32(norm.weight), p32(norm.bias), norm.eps,
                                          out_dtype=cd if last else torch.float32, pre=pre if j == 0 else 0,
                                          pre_scale=pre_scale if j == 0 else 1.0, pre_bias=pre_bias if j == 0 else 0.0)
                continue
            cols = ops.im2col(x, self.kernel_size, self.stride, self.padding, cd,
                              pre if j == 0 else 0, pre_scale if j == 0 else 1.0, pre_bias if j == 0 else 0.0)
            y = ops.gemm(cols, conv_w_c(conv.weight), p32(conv.bias), out_dtype=torch.float32)
            x = norm(y, out_dtype=cd if last else torch.float32, act=ops.ACT_GELU)
        fin = self.encoder[3 * self.num_layers]
        return ops.gemm(x, w_c(fin.weight), p32(fin.bias), out_dtype=torch.float32)

    def forward(self, x):
        return as_nchw_view(self.forward_tokens(as_nhwc(x.float())))


import os as _os
_THIN_CONVS = ((1, 4), (4, 16), (16, 64)) if _os.environ.get("MS2_THIN_CONV3", "0") == "1" else ((1, 4), (4, 16))


class CXBlock(nn.Module):
    def __init__(self, dim, kernel_size=7, padding=3, drop_path=0.0, layer_scale_init_value=1e-6, use_dwconv=True):
        super().__init__()
        assert kernel_size == 7 and padding == 3 and use_dwconv, "only the shipped depth-wise 7x7 block is built"
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=kernel_size, padding=padding, groups=dim)
        self.norm = LayerNorm2d(dim, eps=1e-6)
        self.pwconv1 = Linear(dim, 4 * dim)
        self.pwconv2 = Linear(4 * dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None)

    def forward_tokens(self, x):
        cd = compute_dtype()
        C = x.shape[-1]
        d = ops.dwconv7x7(x, p32(self.dwconv.weight).view(C, 49), p32(self.dwconv.bias))
        t = self.norm(d, out_dtype=cd)
        h = self.pwconv1(t, out_dtype=cd, act=ops.ACT_GELU)
        return self.pwconv2(h, out_dtype=torch.float32, residual=x,
                            colscale=None if self.gamma is None else p32(self.gamma))

    def forward(self, x):
        return as_nchw_view(self.forward_tokens(as_nhwc(x.float())))


class Fuser(nn.Module):
    def __init__(self, layer, num_layers, dim=None, input_projection=False):
        super().__init__()
        assert not input_projection, "input_projection is not used by the shipped configs"
        self.proj = nn.Identity()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])

    def forward_tokens(self, x):
        for layer in self.layers:
            x = layer.forward_tokens(x)
        return x

    def forward(self, x):
        return as_nchw_view(self.forward_tokens(as_nhwc(x.float())))


class MemoryEncoder(nn.Module):
    def __init__(self, out_dim, mask_downsampler, fuser, position_encoding, in_dim=256):
        super().__init__()
        self.mask_downsampler = mask_downsampler
        self.pix_feat_proj = nn.Conv2d(in_dim, in_dim, kernel_size=1)
        self.fuser = fuser
        self.position_encoding = position_encoding
        self.out_proj = nn.Identity()
        if out_dim != in_dim:
            self.out_proj = nn.Conv2d(in_dim, out_dim, kernel_size=1)

    def forward_tokens(self, pix_nhwc, mask_nhwc, pre=0, pre_scale=1.0, pre_bias=0.0):
        """pix fp32 [B,h,w,C]; mask fp32 [B,H,W,1] -> memory features fp32 [B,h,w,out_dim]."""
        m = self.mask_downsampler.forward_tokens(mask_nhwc, pre, pre_scale, pre_bias)
        x = ops.gemm(to_compute(pix_nhwc), w_c(self.pix_feat_proj.weight), p32(self.pix_feat_proj.bias),
                     out_dtype=torch.float32, residual=m)
        x = self.fuser.forward_tokens(x)
        if isinstance(self.out_proj, nn.Conv2d):
            x = ops.gemm(to_compute(x), w_c(self.out_proj.weight), p32(self.out_proj.bias), out_dtype=torch.float32)
        return x

    def forward(self, pix_feat, masks, skip_mask_sigmoid=False):
        """Reference signature: NCHW tensors -> {"vision_features", "vision_pos_enc"}."""
        B, _, H, W = masks.shape
        mask_nhwc = masks.float().contiguous().view(B, H, W, 1)
        x = self.forward_tokens(as_nhwc(pix_feat.float()), mask_nhwc, pre=0 if skip_mask_sigmoid else 1)
        feats = as_nchw_view(x)
        return {"vision_features": feats, "vision_pos_enc": [self.position_encoding(feats)]}
