"""Neutral synthetic-data module: seeded weights for a given {name: shape} layout and the
synthetic inputs of BASELINE.json's configs.  No model code lives here; it is
shared by tests/, bench.py (both arms), __graft_entry__.smoke() and the golden generator.
"""

from collections import OrderedDict
import math

import numpy as np
import torch


def _is_ln2d(name):
    parts = name.split(".")
    if "mask_downsampler" in parts and parts[-2].isdigit() and int(parts[-2]) in (1, 4, 7, 10):
        return True
    if "mask_downscaling" in parts and parts[-2] in ("1", "4"):
        return True
    if "output_upscaling" in parts and parts[-2] == "1":
        return True
    return False


def seeded_weights(spec, seed=0, obj_score_bias=4.0, dtype=torch.float32):
    """Deterministic NON-degenerate weights (SURVEY.md §0 finding 3, §8(d) 'Weights'; strengthened:
    every bias / LayerNorm affine is non-trivial so that a dropped term is caught by parity).
    `spec` is an ordered {name: shape}; tensors are drawn in that order from one CPU generator."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    for name, shape in spec.items():
        shape = tuple(shape)
        parts = name.split(".")
        leaf = parts[-1]
        if name.endswith("pos_embed") or name.endswith("pos_embed_window"):
            t = torch.randn(shape, generator=g) * 0.02
        elif name in ("maskmem_tpos_enc", "no_mem_embed", "no_mem_pos_enc", "no_obj_ptr"):
            t = torch.randn(shape, generator=g) * 0.02
        elif name.endswith("positional_encoding_gaussian_matrix"):
            t = torch.randn(shape, generator=g)
        elif leaf == "gamma":
            t = torch.ones(shape)
        elif (len(parts) >= 2 and "norm" in parts[-2]) or _is_ln2d(name):
            if leaf == "weight":
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 2 and ("embed" in name or "token" in name) and "proj" not in name:
            t = torch.randn(shape, generator=g)                    # nn.Embedding default N(0,1)
        elif leaf == "weight":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "output_upscaling" in name:     # ConvTranspose2d k2s2: one tap of `in` channels per output
                fan_in = shape[0]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # bias
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[name] = t.to(dtype)
    k = "sam_mask_decoder.pred_obj_score_head.layers.2.bias"
    if k in sd:
        sd[k] = torch.full((1,), float(obj_score_bias), dtype=dtype)
    return sd


# ------------------------------------------------------------------ inputs
def random_image(size=1024, seed=0):
    """Config 1: uint8 [size,size,3], rng(seed).integers(0,256)."""
    return np.random.default_rng(seed).integers(0, 256, size=(size, size, 3), dtype=np.uint8)


def fundus_images(n=4, size=1024, seed=0):
    """Config 2: REFUGE-shaped synthetic fundus images + one positive click at the disc centre."""
    rng = np.random.default_rng(seed)
    imgs, pts = [], []
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    s = size / 1024.0
    for _ in range(n):
        cx = size / 2 + rng.uniform(-100, 100) * s
        cy = size / 2 + rng.uniform(-100, 100) * s
        r = rng.uniform(150, 250) * s
        d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
        base = 20.0 + 180.0 * np.clip(1.2 - d / (2.2 * r), 0, 1)
        disc = (d < r) * 40.0
        img = np.stack([base + disc, 0.6 * base + 0.8 * disc, 0.3 * base + 0.5 * disc], -1)
        img = img + rng.normal(0, 8, img.shape)
        imgs.append(np.clip(img, 0, 255).astype(np.uint8))
        pts.append(np.array([[cx, cy]], np.float32))
    return imgs, pts


def btcv_volume(n_slices=96, size=1024, seed=1234, n_objects=1):
    """Config 3: BTCV-shaped volume float32 [T,3,S,S] in 0..255 (three equal channels) with
    `n_objects` drifting ellipses, and per-slice tight boxes [T][n_obj] = [x0,y0,x1,y1]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    s = size / 1024.0
    noise = torch.rand((n_slices, 1, max(size // 8, 4), max(size // 8, 4)), generator=g) * 255.0
    bg = torch.nn.functional.interpolate(noise, size=(size, size), mode="bilinear", align_corners=False)[:, 0]
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
    vol = bg * 0.5
    boxes = []
    for t in range(n_slices):
        row = []
        for o in range(n_objects):
            ang = 2 * math.pi * o / max(n_objects, 1)
            off = 0.0 if n_objects == 1 else 300.0 * s
            cx = size / 2 + off * math.cos(ang) + t * s
            cy = size / 2 + off * math.sin(ang) + 0.5 * t * s
            ax = (180.0 if n_objects == 1 else 70.0) * s
            ay = (140.0 if n_objects == 1 else 55.0) * s
            inside = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0
            vol[t] = vol[t] + inside * 60.0
            row.append([cx - ax, cy - ay, cx + ax, cy + ay])
        boxes.append(row)
    vol = vol.clamp(0, 255)
    return vol[:, None].expand(-1, 3, -1, -1).contiguous(), boxes
