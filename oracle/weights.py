"""Checkpoint layout (names + shapes) of the reference SAM2Base state_dict and a
deterministic, NON-degenerate seeded weight generator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The layout restates what `SAM2Base.state_dict()` produces in the reference
(/root/reference/sam2_train/modeling/sam2_base.py:95-250 and the module files it
instantiates); `tests/golden/state_dict_*.json` holds the key/shape list dumped from the
real reference and `tests/test_oracle_golden.py` checks this spec against it.

Why not default init: SURVEY.md §0 finding 3 — zeros pos-embeds, gamma=1e-6 and a
negative object score make every default-init output the constant -1024.
"""

from collections import OrderedDict
import math

import torch

from .config import hiera_blocks


def _lin(S, name, out_f, in_f):
    S[name + ".weight"] = (out_f, in_f)
    S[name + ".bias"] = (out_f,)


def _ln(S, name, c):
    S[name + ".weight"] = (c,)
    S[name + ".bias"] = (c,)


def _conv(S, name, out_c, in_c, k, groups=1):
    S[name + ".weight"] = (out_c, in_c // groups, k, k)
    S[name + ".bias"] = (out_c,)


def _attn(S, name, dim, internal, kv_in=None):
    kv_in = kv_in or dim
    _lin(S, name + ".q_proj", internal, dim)
    _lin(S, name + ".k_proj", internal, kv_in)
    _lin(S, name + ".v_proj", internal, kv_in)
    _lin(S, name + ".out_proj", dim, internal)


def param_spec(cfg):
    """Ordered {name: shape} of every tensor in the reference state_dict (516 for hiera_s)."""
    S = OrderedDict()
    D, M = cfg["d_model"], cfg["mem_dim"]
    S["maskmem_tpos_enc"] = (cfg["num_maskmem"], 1, 1, M)
    S["no_mem_embed"] = (1, 1, D)
    S["no_mem_pos_enc"] = (1, 1, D)
    S["no_obj_ptr"] = (1, D)
    # ---- Hiera trunk (hieradet.py:171-260)
    p = "image_encoder.trunk."
    E = cfg["embed_dim"]
    S[p + "pos_embed"] = (1, E) + tuple(cfg["window_pos_embed_bkg_spatial_size"])
    S[p + "pos_embed_window"] = (1, E, cfg["window_spec"][0], cfg["window_spec"][0])
    _conv(S, p + "patch_embed.proj", E, 3, 7)
    blocks, _ = hiera_blocks(cfg)
    for i, b in enumerate(blocks):
        q = p + f"blocks.{i}."
        _ln(S, q + "norm1", b["dim"])
        _lin(S, q + "attn.qkv", 3 * b["dim_out"], b["dim"])
        _lin(S, q + "attn.proj", b["dim_out"], b["dim_out"])
        _ln(S, q + "norm2", b["dim_out"])
        _lin(S, q + "mlp.layers.0", 4 * b["dim_out"], b["dim_out"])
        _lin(S, q + "mlp.layers.1", b["dim_out"], 4 * b["dim_out"])
        if b["dim"] != b["dim_out"]:
            _lin(S, q + "proj", b["dim_out"], b["dim"])
    # ---- FPN neck (image_encoder.py:45-99)
    for n, ch in enumerate(cfg["backbone_channel_list"]):
        _conv(S, f"image_encoder.neck.convs.{n}.conv", D, ch, 1)
    # ---- mask_downsample (sam2_base.py:108)
    _conv(S, "mask_downsample", 1, 1, 4)
    # ---- memory attention (memory_attention.py:15-56,102-117)
    for l in range(cfg["mem_attn_layers"]):
        q = f"memory_attention.layers.{l}."
        _attn(S, q + "self_attn", D, D)
        _attn(S, q + "cross_attn_image", D, D, kv_in=M)
        _lin(S, q + "linear1", cfg["mem_attn_ffn"], D)
        _lin(S, q + "linear2", D, cfg["mem_attn_ffn"])
        for k in (1, 2, 3):
            _ln(S, q + f"norm{k}", D)
    _ln(S, "memory_attention.norm", D)
    # ---- memory encoder (memory_encoder.py:17-181)
    q = "memory_encoder."
    cin = 1
    for j in range(4):
        cout = cin * 4
        _conv(S, q + f"mask_downsampler.encoder.{3 * j}", cout, cin, 3)
        _ln(S, q + f"mask_downsampler.encoder.{3 * j + 1}", cout)
        cin = cout
    _conv(S, q + "mask_downsampler.encoder.12", D, cin, 1)
    _conv(S, q + "pix_feat_proj", D, D, 1)
    for j in range(2):
        f = q + f"fuser.layers.{j}."
        S[f + "gamma"] = (D,)
        _conv(S, f + "dwconv", D, D, 7, groups=D)
        _ln(S, f + "norm", D)
        _lin(S, f + "pwconv1", 4 * D, D)
        _lin(S, f + "pwconv2", D, 4 * D)
    _conv(S, q + "out_proj", M, D, 1)
    # ---- prompt encoder (prompt_encoder.py:17-66)
    q = "sam_prompt_encoder."
    S[q + "pe_layer.positional_encoding_gaussian_matrix"] = (2, D // 2)
    for j in range(4):
        S[q + f"point_embeddings.{j}.weight"] = (1, D)
    S[q + "not_a_point_embed.weight"] = (1, D)
    _conv(S, q + "mask_downscaling.0", 4, 1, 2)
    _ln(S, q + "mask_downscaling.1", 4)
    _conv(S, q + "mask_downscaling.3", 16, 4, 2)
    _ln(S, q + "mask_downscaling.4", 16)
    _conv(S, q + "mask_downscaling.6", D, 16, 1)
    S[q + "no_mask_embed.weight"] = (1, D)
    # ---- mask decoder (mask_decoder.py:13-108, transformer.py:28-163)
    q = "sam_mask_decoder."
    for l in range(2):
        t = q + f"transformer.layers.{l}."
        _attn(S, t + "self_attn", D, D)
        _ln(S, t + "norm1", D)
        _attn(S, t + "cross_attn_token_to_image", D, D // 2)
        _ln(S, t + "norm2", D)
        _lin(S, t + "mlp.layers.0", 2048, D)
        _lin(S, t + "mlp.layers.1", D, 2048)
        _ln(S, t + "norm3", D)
        _ln(S, t + "norm4", D)
        _attn(S, t + "cross_attn_image_to_token", D, D // 2)
    _attn(S, q + "transformer.final_attn_token_to_image", D, D // 2)
    _ln(S, q + "transformer.norm_final_attn", D)
    S[q + "iou_token.weight"] = (1, D)
    S[q + "mask_tokens.weight"] = (4, D)
    S[q + "obj_score_token.weight"] = (1, D)
    S[q + "output_upscaling.0.weight"] = (D, D // 4, 2, 2)   # ConvTranspose2d: [in, out, k, k]
    S[q + "output_upscaling.0.bias"] = (D // 4,)
    _ln(S, q + "output_upscaling.1", D // 4)
    S[q + "output_upscaling.3.weight"] = (D // 4, D // 8, 2, 2)
    S[q + "output_upscaling.3.bias"] = (D // 8,)
    _conv(S, q + "conv_s0", D // 8, D, 1)
    _conv(S, q + "conv_s1", D // 4, D, 1)
    for j in range(4):
        h = q + f"output_hypernetworks_mlps.{j}.layers."
        _lin(S, h + "0", D, D)
        _lin(S, h + "1", D, D)
        _lin(S, h + "2", D // 8, D)
    h = q + "iou_prediction_head.layers."
    _lin(S, h + "0", D, D)
    _lin(S, h + "1", D, D)
    _lin(S, h + "2", 4, D)
    h = q + "pred_obj_score_head.layers."
    _lin(S, h + "0", D, D)
    _lin(S, h + "1", D, D)
    _lin(S, h + "2", 1, D)
    # ---- object pointer projection (sam2_base.py:236-244)
    for j in range(3):
        _lin(S, f"obj_ptr_proj.layers.{j}", D, D)
    return S


def make_state_dict(cfg, seed=0, obj_score_bias=4.0, dtype=torch.float32):
    """Seeded non-degenerate weights for the reference layout (generator lives in the neutral
    root module `synth_data.py` so that bench.py's product arm can use it without importing oracle/)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from synth_data import seeded_weights
    return seeded_weights(param_spec(cfg), seed=seed, obj_score_bias=obj_score_bias, dtype=dtype)
