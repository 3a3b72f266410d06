"""Input-independent position tables.  The reference recomputes them on every call
(position_encoding.py:79-112,138-149,167-185); here they are built once per (shape, device) with
plain tensor math and cached — they are constants, not part of the per-slice work."""
import math

import torch
from torch import nn

_TABLES = {}


def sine_table(num_pos_feats, H, W, device, temperature=10000.0):
    """PositionEmbeddingSine (normalize=True, scale=2*pi) as a token-major fp32 table [H,W,C]."""
    key = ("sine", num_pos_feats, H, W, str(device), temperature)
    if key not in _TABLES:
        npf = num_pos_feats // 2
        y = torch.arange(1, H + 1, dtype=torch.float32, device=device)
        x = torch.arange(1, W + 1, dtype=torch.float32, device=device)
        y = y / (y[-1] + 1e-6) * (2 * math.pi)
        x = x / (x[-1] + 1e-6) * (2 * math.pi)
        dim_t = torch.arange(npf, dtype=torch.float32, device=device)
        dim_t = temperature ** (2 * (dim_t // 2) / npf)
        px, py = x[:, None] / dim_t, y[:, None] / dim_t
        px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
        py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
        _TABLES[key] = torch.cat((py[:, None, :].expand(H, W, npf), px[None, :, :].expand(H, W, npf)), dim=2).contiguous()
    return _TABLES[key]


def rope_table(D, side_w, side_h, theta, device):
    """axial RoPE angles -> (cos, sin) fp32 [side_w*side_h, D/2] (compute_axial_cis)."""
    key = ("rope", D, side_w, side_h, float(theta), str(device))
    if key not in _TABLES:
        freqs = 1.0 / (theta ** (torch.arange(0, D, 4)[: D // 4].float() / D))
        t = torch.arange(side_w * side_h, dtype=torch.float32)
        ang = torch.cat([torch.outer((t % side_w).float(), freqs),
                         torch.outer(torch.div(t, side_w, rounding_mode="floor").float(), freqs)], dim=-1)
        _TABLES[key] = (ang.cos().contiguous().to(device), ang.sin().contiguous().to(device))
    return _TABLES[key]


class PositionEmbeddingSine(nn.Module):
    """Parameter-free; kept as a module so the YAML tree instantiates unchanged."""

    def __init__(self, num_pos_feats, temperature=10000, normalize=True, scale=None):
        super().__init__()
        assert num_pos_feats % 2 == 0, "Expecting even model width"
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        assert normalize and scale in (None, 2 * math.pi), "only the shipped configuration is built"
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature

    def table(self, H, W, device):
        return sine_table(self.num_pos_feats, H, W, device, float(self.temperature))

    @torch.no_grad()
    def forward(self, x):
        """x: [B,C,H,W] -> [B,num_pos_feats,H,W] (expanded view of the cached table)."""
        t = self.table(x.shape[-2], x.shape[-1], x.device)
        return t.permute(2, 0, 1)[None].expand(x.shape[0], -1, -1, -1)
