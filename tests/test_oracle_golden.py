"""CPU: pin the oracle (oracle/sam2_oracle.py) to known answers produced by the REAL reference
(tests/golden/make_golden.py). fp32; tolerances are abs and written per check."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.config import get_config
from oracle.weights import param_spec, make_state_dict
from oracle.sam2_oracle import (OracleSAM2, OracleImagePredictor, OracleVideoPredictor,
                                connected_components_np)
from synth_data import random_image, fundus_images, btcv_volume

G = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, tol, what):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    assert err <= tol, f"{what}: max abs err {err} > {tol}"


@pytest.mark.parametrize("cfg", ["sam2_hiera_s", "sam2_hiera_t"])
def test_checkpoint_layout_matches_reference(cfg):
    ref = json.load(open(f"{G}/state_dict_{cfg}.json"))
    spec = param_spec(get_config(cfg))
    assert set(spec) == set(ref)
    for k, shp in spec.items():
        assert list(shp) == ref[k], k


def test_cc_oracle_matches_kernel_transliteration():
    z = np.load(f"{G}/cc_cases.npz")
    for i in range(int(z["n"])):
        m = z[f"mask_{i}"]
        l, c = connected_components_np(m[None, None])
        assert np.array_equal(l[0, 0], z[f"labels_{i}"]), i
        assert np.array_equal(c[0, 0], z[f"counts_{i}"]), i


def test_modules_match_reference():
    z = np.load(f"{G}/modules_hiera_t.npz")
    cfg = get_config("sam2_hiera_t", image_size=512)
    m = OracleSAM2(cfg, make_state_dict(cfg))
    g = torch.Generator().manual_seed(7)
    B, HW = 2, 32 * 32
    curr = torch.randn(HW, B, 256, generator=g)
    curr_pos = torch.randn(HW, B, 256, generator=g)
    Lk = 2 * HW + 8
    memory = torch.randn(Lk, B, 64, generator=g)
    memory_pos = torch.randn(Lk, B, 64, generator=g)
    y = m.memory_attention([curr], [curr_pos], memory, memory_pos, 8)
    _close(y[::4], z["memattn_out_sub"], 2e-4, "memory_attention")
    pix = torch.randn(HW, B, 256, generator=g)
    hi = torch.randn(B, 1, 512, 512, generator=g) * 3
    for flag in (False, True):
        f, pe = m.encode_new_memory([pix], [(32, 32)], hi, flag)
        _close(f, z[f"memenc_feat_{int(flag)}"], 2e-4, f"memory_encoder {flag}")
    _close(pe[0][0], z["memenc_pos"], 1e-6, "memenc pos")
    emb = torch.randn(B, 256, 32, 32, generator=g)
    hr0 = torch.randn(B, 32, 128, 128, generator=g)
    hr1 = torch.randn(B, 64, 64, 64, generator=g)
    pts = {"point_coords": torch.tensor([[[100.0, 200.0], [300.0, 50.0]], [[10.0, 20.0], [400.0, 500.0]]]),
           "point_labels": torch.tensor([[1, 0], [2, 3]], dtype=torch.int32)}
    for mm in (False, True):
        r = m.forward_sam_heads(emb, point_inputs=pts, high_res_features=[hr0, hr1], multimask_output=mm)
        _close(r[0], z[f"heads_low_{int(mm)}"], 5e-4, "heads low")
        _close(r[2], z[f"heads_ious_{int(mm)}"], 1e-5, "heads ious")
        _close(r[5], z[f"heads_ptr_{int(mm)}"], 1e-4, "heads ptr")
        _close(r[6], z[f"heads_obj_{int(mm)}"], 1e-4, "heads obj")
    mask_in = (torch.rand(B, 1, 512, 512, generator=g) > 0.5).float()
    r = m.use_mask_as_output(emb, [hr0, hr1], mask_in)
    _close(r[0], z["maskout_low"], 1e-5, "mask-as-output low")
    _close(r[5], z["maskout_ptr"], 1e-4, "mask-as-output ptr")
    _close(m.get_dense_pe(), z["dense_pe"], 1e-5, "dense pe")


def test_image_predictor_hiera_t_config1():
    z = np.load(f"{G}/image_hiera_t_1024.npz")
    cfg = get_config("sam2_hiera_t")
    m = OracleSAM2(cfg, make_state_dict(cfg))
    p = OracleImagePredictor(m)
    p.set_image(random_image(1024, 0))
    _close(p._features["image_embed"][..., ::4, ::4], z["image_embed_sub"], 5e-4, "image_embed")
    _close(p._features["high_res_feats"][0][..., ::16, ::16], z["high_res0_sub"], 5e-4, "hr0")
    _close(p._features["high_res_feats"][1][..., ::8, ::8], z["high_res1_sub"], 5e-4, "hr1")
    masks, ious, low = p.predict(point_coords=np.array([[512, 512]]), point_labels=np.array([1]),
                                 multimask_output=True, return_logits=True)
    _close(low, z["low_res"], 1e-4, "low_res logits")
    _close(ious, z["ious"], 1e-5, "ious")
    _close(masks[:, ::4, ::4], z["masks_sub"], 1e-4, "masks")
    _, ious1, low1 = p.predict(box=np.array([300, 350, 700, 800]), multimask_output=False, return_logits=True)
    _close(low1, z["box_low_res"], 1e-4, "box low_res")
    _close(ious1, z["box_ious"], 1e-5, "box ious")


def test_image_predictor_hiera_s_batch():
    z = np.load(f"{G}/image_hiera_s_1024.npz")
    cfg = get_config("sam2_hiera_s")
    m = OracleSAM2(cfg, make_state_dict(cfg))
    p = OracleImagePredictor(m)
    imgs, pts = fundus_images(4, 1024, 0)
    p.set_image_batch(imgs)
    _close(p._features["image_embed"][..., ::4, ::4], z["image_embed_sub"], 5e-4, "image_embed")
    masks, ious, low = p.predict_batch(pts, [np.array([1])] * 4, multimask_output=True, return_logits=True)
    _close(np.stack(low), z["low_res"], 1e-4, "low_res")
    _close(np.stack(ious), z["ious"], 1e-5, "ious")


def _video(cfg_name, size, n_slices, n_obj, prompt_frames, absent, seed):
    cfg = get_config(cfg_name, image_size=size)
    m = OracleSAM2(cfg, make_state_dict(cfg))
    vp = OracleVideoPredictor(m)
    vol, boxes = btcv_volume(n_slices, size, seed, n_obj)
    st = vp.init_state(vol, size, size)
    for f in prompt_frames:
        for o in range(n_obj):
            if (f, o) in absent:
                vp.add_new_mask(st, f, o + 1, torch.zeros(size, size))
            else:
                vp.add_new_bbox(st, f, o + 1, boxes[f][o], clear_old_points=False)
    outs = {f: mk for f, _, mk in vp.propagate_in_video(st, start_frame_idx=0)}
    return st, outs


@pytest.mark.parametrize("case", ["s1", "t2"])
def test_video_predictor_matches_reference(case):
    if case == "s1":
        args, fname = ("sam2_hiera_s", 512, 7, 1, (0, 2, 4), (), 1234), "video_hiera_s_512.npz"
    else:
        args, fname = ("sam2_hiera_t", 512, 6, 2, (0, 3), ((3, 1),), 77), "video_hiera_t_512_2obj.npz"
    z = np.load(f"{G}/{fname}")
    st, outs = _video(*args)
    od = st["output_dict"]
    for f in range(args[2]):
        o = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)
        _close(o["pred_masks"], z[f"pred_masks_{f}"], 2e-3, f"pred_masks frame {f}")
        _close(o["obj_ptr"], z[f"obj_ptr_{f}"], 1e-3, f"obj_ptr frame {f}")
        _close(o["maskmem_features"][..., ::4, ::4], z[f"maskmem_sub_{f}"], 1e-3, f"maskmem frame {f}")
        _close(outs[f][..., ::4, ::4], z["video_res_masks_sub"][f], 2e-3, f"video masks frame {f}")


def test_video_clicks_reverse_reset_match_reference():
    """click prompts + refinement click + forward/reverse propagation + reset_state + second session
    (golden: video_clicks_hiera_t_512.npz, made by make_golden.py video_clicks on the real reference)."""
    z = np.load(f"{G}/video_clicks_hiera_t_512.npz")
    size, T = 512, 6
    cfg = get_config("sam2_hiera_t", image_size=size)
    vp = OracleVideoPredictor(OracleSAM2(cfg, make_state_dict(cfg)))
    vol, boxes = btcv_volume(T, size, 55, 1)
    st = vp.init_state(vol, size, size)

    def centre(f):
        x0, y0, x1, y1 = boxes[f][0]
        return [(x0 + x1) / 2.0, (y0 + y1) / 2.0]
    c2 = centre(2)
    _, _, vr = vp.add_new_points(st, 2, 1, [c2], [1], clear_old_points=False)
    _close(vr[..., ::4, ::4], z["a/click1_video_res_sub"], 2e-3, "first click")
    _, _, vr = vp.add_new_points(st, 2, 1, [[c2[0] + 150.0, c2[1] + 120.0]], [0], clear_old_points=False)
    _close(vr[..., ::4, ::4], z["a/click2_video_res_sub"], 2e-3, "refinement click")
    fwd = {f: mk for f, _, mk in vp.propagate_in_video(st, start_frame_idx=2)}
    rev = {f: mk for f, _, mk in vp.propagate_in_video(st, start_frame_idx=2, reverse=True)}
    assert sorted(fwd) == list(z["a/fwd_frames"]) and sorted(rev) == list(z["a/rev_frames"])
    od = st["output_dict"]
    for f in range(T):
        o = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)
        _close(o["pred_masks"], z[f"a/pred_masks_{f}"], 2e-3, f"a pred_masks {f}")
        _close(o["obj_ptr"], z[f"a/obj_ptr_{f}"], 1e-3, f"a obj_ptr {f}")
        _close((fwd[f] if f in fwd else rev[f])[..., ::4, ::4], z["a/video_res_sub"][f], 2e-3, f"a video res {f}")
    vp.reset_state(st)
    vp.add_new_bbox(st, 0, 7, boxes[0][0], clear_old_points=False)
    vp.add_new_points(st, 3, 7, [centre(3)], [1], clear_old_points=False)
    b = {f: mk for f, _, mk in vp.propagate_in_video(st, start_frame_idx=0, max_frame_num_to_track=4)}
    assert sorted(b) == list(z["b/frames"])
    for i, f in enumerate(sorted(b)):
        o = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)
        _close(o["pred_masks"], z[f"b/pred_masks_{f}"], 2e-3, f"b pred_masks {f}")
        _close(b[f][..., ::4, ::4], z["b/video_res_sub"][i], 2e-3, f"b video res {f}")


def test_postprocess_and_non_overlap_match_reference():
    from oracle.sam2_oracle import apply_non_overlapping_constraints, postprocess_masks
    z = np.load(f"{G}/postprocess_cases.npz")
    masks = torch.from_numpy(z["pp/masks"])
    for tag, (ha, sa, thr) in {"h8_s4": (8.0, 4.0, 0.0), "h8_s0": (8.0, 0.0, 0.0), "h3_s6_t05": (3.0, 6.0, 0.5)}.items():
        got = postprocess_masks(masks.clone(), (96, 80), thr, ha, sa)
        _close(got, z[f"pp/{tag}"], 1e-6, f"postprocess {tag}")
    for n in (1, 2, 5):
        got = apply_non_overlapping_constraints(torch.from_numpy(z[f"no/in_{n}"]))
        assert np.array_equal(got.numpy(), z[f"no/out_{n}"]), n


def test_ingest_jpeg_matches_reference(tmp_path):
    from oracle.sam2_oracle import load_video_frames
    z = np.load(f"{G}/ingest_jpeg.npz")
    for i in range(5):
        (tmp_path / f"{i}.jpg").write_bytes(z[f"jpeg_{i}"].tobytes())
    frames, h, w = load_video_frames(str(tmp_path), 64)
    assert (h, w) == tuple(z["hw"])
    _close(frames, z["sync"], 1e-6, "sync loader")
    _close(frames, z["async"], 1e-6, "async loader")
