"""ImageEncoder = Hiera trunk + FPN neck (reference backbones/image_encoder.py:13-133), token-major.

The 1x1 lateral convs are GEMMs over tokens; the nearest-x2 top-down add is one kernel; the sine
position encodings are cached tables (the reference materialises 84 MiB of them per frame).
Outputs keep the reference's NCHW *shape* as channels-last strided views, so reference callers work
unchanged while our own consumers get the NHWC memory back for free.
"""
import torch
from torch import nn

from ... import ops
from ...runtime import CACHE, compute_dtype, p32, w_c
from ..sam2_utils import as_nchw_view, as_nhwc, to_compute


class FpnNeck(nn.Module):
    def __init__(self, position_encoding, d_model, backbone_channel_list, kernel_size=1, stride=1, padding=0,
                 fpn_interp_model="bilinear", fuse_type="sum", fpn_top_down_levels=None):
        super().__init__()
        assert kernel_size == 1 and stride == 1 and padding == 0
        assert fuse_type == "sum" and fpn_interp_model == "nearest", "only the shipped FPN variant is built"
        self.position_encoding = position_encoding
        self.backbone_channel_list = list(backbone_channel_list)
        self.d_model = d_model
        self.convs = nn.ModuleList()
        for dim in backbone_channel_list:
            current = nn.Sequential()
            current.add_module("conv", nn.Conv2d(dim, d_model, kernel_size=1))
            self.convs.append(current)
        if fpn_top_down_levels is None:
            fpn_top_down_levels = range(len(self.convs))
        self.fpn_top_down_levels = list(fpn_top_down_levels)

    def lateral(self, i, x_nhwc, fold=None):
        """1x1 conv of level i as a token GEMM.  `fold` = a second 1x1 conv (nn.Conv2d) applied right after it on a
        level that receives no top-down feature (SAM2Base.forward_image: conv_s0 / conv_s1 on the two high-resolution
        levels, sam2_base.py:464-476): two 1x1 convs in a row are ONE linear map, W = W2 W1, b = W2 b1 + b2 (folded in
        fp32), so the 256-channel intermediate (512 MiB per 8-slice batch at level 0) is never written, cast or read."""
        n = len(self.convs) - 1
        conv = self.convs[n - i].conv
        if fold is None:
            return ops.gemm(to_compute(x_nhwc), w_c(conv.weight), p32(conv.bias), out_dtype=torch.float32)
        dt = compute_dtype()

        def make(w1, b1, w2, b2):
            W2, W1 = w2.detach().float().reshape(w2.shape[0], -1), w1.detach().float().reshape(w1.shape[0], -1)
            return ((W2 @ W1).to(dt).contiguous(), (W2 @ b1.detach().float() + b2.detach().float()).contiguous())
        W, b = CACHE.get((conv.weight, conv.bias, fold.weight, fold.bias), ("fold1x1", dt), make)
        return ops.gemm(to_compute(x_nhwc), W, b, out_dtype=torch.float32)

    def forward_tokens(self, xs, fold=None):
        """`fold`: {level: nn.Conv2d} of 1x1 convs to fold into the lateral conv of levels without top-down input."""
        n = len(self.convs) - 1
        out = [None] * len(self.convs)
        prev = None
        for i in range(n, -1, -1):
            f = None if fold is None else fold.get(i)
            if f is not None:
                assert i not in self.fpn_top_down_levels, "only levels without a top-down sum can be folded"
            lat = self.lateral(i, xs[i], f)
            if i in self.fpn_top_down_levels and prev is not None:
                ops.upsample2x_add_(lat, prev)
            prev = lat
            out[i] = lat
        return out

    def forward(self, xs):
        feats = self.forward_tokens([as_nhwc(x) for x in xs])
        out = [as_nchw_view(f) for f in feats]
        pos = [self.position_encoding(o) for o in out]
        return out, pos


class ImageEncoder(nn.Module):
    def __init__(self, trunk, neck, scalp=0):
        super().__init__()
        self.trunk, self.neck, self.scalp = trunk, neck, scalp
        assert self.trunk.channel_list == self.neck.backbone_channel_list, (
            f"Channel dims of trunk and neck do not match. Trunk: {self.trunk.channel_list}, "
            f"neck: {self.neck.backbone_channel_list}")

    def forward_tokens(self, sample, fold=None):
        feats = self.neck.forward_tokens(self.trunk.forward_tokens(sample), fold)
        if self.scalp > 0:
            feats = feats[: -self.scalp]
        return feats

    def forward(self, sample):
        feats = [as_nchw_view(f) for f in self.forward_tokens(sample)]
        pos = [self.neck.position_encoding(f) for f in feats]
        return {"vision_features": feats[-1], "vision_pos_enc": pos, "backbone_fpn": feats}
