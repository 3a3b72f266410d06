"""ImageEncoder = Hiera trunk + FPN neck (reference backbones/image_encoder.py:13-133), token-major.

The 1x1 lateral convs are GEMMs over tokens; the nearest-x2 top-down add is one kernel; the sine
position encodings are cached tables (the reference materialises 84 MiB of them per frame).
Outputs keep the reference's NCHW *shape* as channels-last strided views, so reference callers work
unchanged while our own consumers get the NHWC memory back for free.
"""
import torch
from torch import nn

from ... import ops
from ...runtime import compute_dtype, p32, w_c
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
        """1x1 conv of level i as a token GEMM; `fold=(W2,b2)` applies a second 1x1 conv folded in."""
        n = len(self.convs) - 1
        conv = self.convs[n - i].conv
        return ops.gemm(to_compute(x_nhwc), w_c(conv.weight), p32(conv.bias), out_dtype=torch.float32)

    def forward_tokens(self, xs):
        n = len(self.convs) - 1
        out = [None] * len(self.convs)
        prev = None
        for i in range(n, -1, -1):
            lat = self.lateral(i, xs[i])
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

    def forward_tokens(self, sample):
        feats = self.neck.forward_tokens(self.trunk.forward_tokens(sample))
        if self.scalp > 0:
            feats = feats[: -self.scalp]
        return feats

    def forward(self, sample):
        feats = [as_nchw_view(f) for f in self.forward_tokens(sample)]
        pos = [self.neck.position_encoding(f) for f in feats]
        return {"vision_features": feats[-1], "vision_pos_enc": pos, "backbone_fpn": feats}
