"""Model hyper-parameters of the two shipped reference configs, as plain dicts.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/sam2_train/sam2_hiera_s.yaml:1-116 and sam2_hiera_t.yaml:1-118
plus the post-processing overrides of build_sam.py:24-30 (image) / :51-66 (video).
"""

import copy

_COMMON = dict(
    embed_dim=96,
    num_heads=1,
    window_spec=(8, 4, 14, 7),          # hieradet.py:188-193 (default, not in YAML)
    window_pos_embed_bkg_spatial_size=(7, 7),
    q_pool=3,
    d_model=256,
    backbone_channel_list=(768, 384, 192, 96),
    fpn_top_down_levels=(2, 3),
    scalp=1,
    mem_dim=64,
    num_maskmem=7,
    image_size=1024,
    backbone_stride=16,
    mem_attn_layers=4,
    mem_attn_ffn=2048,
    rope_theta=10000.0,
    sigmoid_scale_for_mem_enc=20.0,
    sigmoid_bias_for_mem_enc=-10.0,
    max_obj_ptrs_in_encoder=16,
    max_cond_frames_in_attn=-1,
    memory_temporal_stride_for_eval=1,
    multimask_min_pt_num=0,
    multimask_max_pt_num=1,
    # build_sam.py overrides
    dynamic_multimask_via_stability=True,
    dynamic_multimask_stability_delta=0.05,
    dynamic_multimask_stability_thresh=0.98,
    binarize_mask_from_pts_for_mem_enc=True,   # video predictor only (build_sam.py:62)
    fill_hole_area=8,                          # video predictor only (build_sam.py:64)
)

CONFIGS = {
    "sam2_hiera_s": dict(_COMMON, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13)),
    "sam2_hiera_t": dict(_COMMON, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9)),
}


def get_config(name, **overrides):
    name = name.replace(".yaml", "")
    cfg = copy.deepcopy(CONFIGS[name])
    cfg["name"] = name
    cfg.update(overrides)
    return cfg


def hiera_blocks(cfg):
    """Per-block (dim_in, dim_out, heads, window, q_pool) table — hieradet.py:196-260."""
    stages = cfg["stages"]
    depth = sum(stages)
    stage_ends = [sum(stages[:i]) - 1 for i in range(1, len(stages) + 1)]
    q_pool_blocks = [x + 1 for x in stage_ends[:-1]][: cfg["q_pool"]]
    embed_dim, heads = cfg["embed_dim"], cfg["num_heads"]
    cur_stage = 1
    out = []
    for i in range(depth):
        dim_out = embed_dim
        window = cfg["window_spec"][cur_stage - 1]
        if i in cfg["global_att_blocks"]:
            window = 0
        if i - 1 in stage_ends:
            dim_out = embed_dim * 2
            heads = heads * 2
            cur_stage += 1
        out.append(dict(dim=embed_dim, dim_out=dim_out, heads=heads, window=window,
                        q_pool=(i in q_pool_blocks)))
        embed_dim = dim_out
    return out, stage_ends
