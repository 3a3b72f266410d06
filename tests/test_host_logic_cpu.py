"""CPU: the product's HOST logic (module tree, layout plumbing, predictors' state machines, checkpoint
handling) checked end-to-end against the golden fixtures of the real reference, with the native ops
replaced by their plain-torch statements from tests/ref_ops.py (a test-only fake backend; the product
package has no CPU path of its own)."""
import os

import numpy as np
import pytest
import torch

import ref_ops
from oracle.config import get_config
from oracle.weights import make_state_dict
from synth_data import btcv_volume, fundus_images, random_image

G = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, tol, what):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, np.float32)
    b = np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    assert err <= tol, f"{what}: max abs err {err} > {tol}"


@pytest.fixture()
def fake_backend(monkeypatch):
    import medsam2_b200
    ref_ops.install(monkeypatch)
    with medsam2_b200.compute(torch.float32):
        yield


def _build(cfg, video, image_size=1024):
    import medsam2_b200
    fn = medsam2_b200.build_sam2_video_predictor if video else medsam2_b200.build_sam2
    m = fn(cfg, device="cpu", hydra_overrides_extra=[f"++model.image_size={image_size}"])
    m.load_state_dict(make_state_dict(get_config(cfg)), strict=True)
    return m


def test_image_predictor_config1(fake_backend):
    from medsam2_b200 import SAM2ImagePredictor
    z = np.load(f"{G}/image_hiera_t_1024.npz")
    p = SAM2ImagePredictor(_build("sam2_hiera_t", video=False))
    p.set_image(random_image(1024, 0))
    _close(p._features["image_embed"][..., ::4, ::4], z["image_embed_sub"], 5e-4, "image_embed")
    _close(p._features["high_res_feats"][0][..., ::16, ::16], z["high_res0_sub"], 5e-4, "hr0")
    _close(p._features["high_res_feats"][1][..., ::8, ::8], z["high_res1_sub"], 5e-4, "hr1")
    masks, ious, low = p.predict(point_coords=np.array([[512, 512]]), point_labels=np.array([1]),
                                 multimask_output=True, return_logits=True)
    _close(low, z["low_res"], 2e-4, "low_res")
    _close(ious, z["ious"], 1e-5, "ious")
    _close(masks[:, ::4, ::4], z["masks_sub"], 2e-4, "masks")
    _, ious1, low1 = p.predict(box=np.array([300, 350, 700, 800]), multimask_output=False, return_logits=True)
    _close(low1, z["box_low_res"], 2e-4, "box low_res")
    _close(ious1, z["box_ious"], 1e-5, "box ious")


def test_image_predictor_batch(fake_backend):
    from medsam2_b200 import SAM2ImagePredictor
    z = np.load(f"{G}/image_hiera_s_1024.npz")
    p = SAM2ImagePredictor(_build("sam2_hiera_s", video=False))
    imgs, pts = fundus_images(2, 1024, 0)
    p.set_image_batch(imgs)
    _close(p._features["image_embed"][..., ::4, ::4], z["image_embed_sub"][:2], 5e-4, "image_embed")
    masks, ious, low = p.predict_batch(pts, [np.array([1])] * 2, multimask_output=True, return_logits=True)
    _close(np.stack(low), z["low_res"][:2], 2e-4, "low_res")
    _close(np.stack(ious), z["ious"][:2], 1e-5, "ious")


@pytest.mark.parametrize("case", ["s1", "t2"])
def test_video_predictor(fake_backend, case):
    if case == "s1":
        cfg, size, T, n_obj, prompts, absent, seed, fname = "sam2_hiera_s", 512, 7, 1, (0, 2, 4), (), 1234, "video_hiera_s_512.npz"
    else:
        cfg, size, T, n_obj, prompts, absent, seed, fname = "sam2_hiera_t", 512, 6, 2, (0, 3), ((3, 1),), 77, "video_hiera_t_512_2obj.npz"
    z = np.load(f"{G}/{fname}")
    m = _build(cfg, video=True, image_size=size)
    # batched slice encoding + preflight memory encoder over several prompted frames per pass (golden answers are from
    # the reference's frame-at-a-time flow)
    m.feature_cache_size, m.feature_encode_batch = 8, 4
    vol, boxes = btcv_volume(T, size, seed, n_obj)
    st = m.val_init_state(imgs_tensor=vol, video_height=size, video_width=size)
    for f in prompts:
        for o in range(n_obj):
            if (f, o) in absent:
                m.train_add_new_mask(inference_state=st, frame_idx=f, obj_id=o + 1, mask=torch.zeros(size, size))
            else:
                m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=o + 1, bbox=torch.tensor(boxes[f][o]),
                                     clear_old_points=False)
    outs = {f: mk.clone() for f, _, mk in m.propagate_in_video(st, start_frame_idx=0)}
    od = st["output_dict"]
    for f in range(T):
        o = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"].get(f)
        _close(o["pred_masks"], z[f"pred_masks_{f}"], 3e-3, f"pred_masks frame {f}")
        _close(o["obj_ptr"], z[f"obj_ptr_{f}"], 1e-3, f"obj_ptr frame {f}")
        _close(o["maskmem_features"][..., ::4, ::4], z[f"maskmem_sub_{f}"], 1e-3, f"maskmem frame {f}")
        _close(outs[f][..., ::4, ::4], z["video_res_masks_sub"][f], 3e-3, f"video masks frame {f}")


def test_submodule_surface(fake_backend):
    """func_2d-style direct calls against module-level golden answers."""
    z = np.load(f"{G}/modules_hiera_t.npz")
    m = _build("sam2_hiera_t", video=True, image_size=512)
    g = torch.Generator().manual_seed(7)
    B, HW = 2, 32 * 32
    curr = torch.randn(HW, B, 256, generator=g)
    curr_pos = torch.randn(HW, B, 256, generator=g)
    Lk = 2 * HW + 8
    memory = torch.randn(Lk, B, 64, generator=g)
    memory_pos = torch.randn(Lk, B, 64, generator=g)
    y = m.memory_attention(curr=[curr], curr_pos=[curr_pos], memory=memory, memory_pos=memory_pos, num_obj_ptr_tokens=8)
    _close(y[::4], z["memattn_out_sub"], 3e-4, "memory_attention")
    pix = torch.randn(HW, B, 256, generator=g)
    hi = torch.randn(B, 1, 512, 512, generator=g) * 3
    for flag in (False, True):
        f, pe = m._encode_new_memory([pix], [(32, 32)], hi, is_mask_from_pts=flag)
        _close(f, z[f"memenc_feat_{int(flag)}"], 3e-4, f"memory_encoder {flag}")
    _close(pe[0][0], z["memenc_pos"], 1e-6, "memenc pos")
    emb = torch.randn(B, 256, 32, 32, generator=g)
    hr0 = torch.randn(B, 32, 128, 128, generator=g)
    hr1 = torch.randn(B, 64, 64, 64, generator=g)
    pts = {"point_coords": torch.tensor([[[100.0, 200.0], [300.0, 50.0]], [[10.0, 20.0], [400.0, 500.0]]]),
           "point_labels": torch.tensor([[1, 0], [2, 3]], dtype=torch.int32)}
    for mm in (False, True):
        r = m._forward_sam_heads(emb, point_inputs=pts, high_res_features=[hr0, hr1], multimask_output=mm)
        _close(r[0], z[f"heads_low_{int(mm)}"], 5e-4, "heads low")
        _close(r[2], z[f"heads_ious_{int(mm)}"], 1e-5, "heads ious")
        _close(r[5], z[f"heads_ptr_{int(mm)}"], 1e-4, "heads ptr")
        _close(r[6], z[f"heads_obj_{int(mm)}"], 1e-4, "heads obj")
    mask_in = (torch.rand(B, 1, 512, 512, generator=g) > 0.5).float()
    r = m._use_mask_as_output(emb, [hr0, hr1], mask_in)
    _close(r[0], z["maskout_low"], 1e-5, "mask-as-output low")
    _close(r[5], z["maskout_ptr"], 1e-4, "mask-as-output ptr")
    _close(m.sam_prompt_encoder.get_dense_pe(), z["dense_pe"], 1e-5, "dense pe")


@pytest.mark.parametrize("tag,shape", [("b1_p3", (1, 3, 3)), ("b2_p2", (2, 3, 2))])
def test_func2d_memory_bank_step(fake_backend, tag, shape):
    """The 2D memory-bank validation step (func_2d/function.py:423-534) replayed on the product
    by tests/func2d_replay.py against the answers of the real reference: memory attention without object pointers,
    bank sampling of shape [B,B], many prompts per image through `cell_nums`, memory encoding of the merged mask."""
    from func2d_replay import make_inputs, replay
    z = np.load(f"{G}/func2d_hiera_t_512.npz")
    m = _build("sam2_hiera_t", video=True, image_size=512)
    with torch.no_grad():
        r = replay(m, *make_inputs(*shape), sampled_indices=torch.from_numpy(z[f"{tag}/sampled_indices"]))
    _close(r["similarity"], z[f"{tag}/similarity"], 1e-5, "similarity")
    _close(r["memattn"][::4], z[f"{tag}/memattn_sub"], 5e-4, "memory attention")
    _close(r["low_res"], z[f"{tag}/low_res"], 1e-3, "low-res masks")
    _close(r["iou"], z[f"{tag}/iou"], 1e-4, "iou")
    _close(r["obj"], z[f"{tag}/obj"], 1e-4, "object score")
    _close(r["maskmem_feat"][..., ::2, ::2], z[f"{tag}/maskmem_feat_sub"], 5e-4, "maskmem features")
    _close(r["maskmem_pos"][..., ::2, ::2], z[f"{tag}/maskmem_pos_sub"], 1e-6, "maskmem pos")


def test_validation_driver(fake_backend):
    """`validation_sam` (reference func_3d/function.py:198-314): two prompted volumes + one without annotations.  The
    driver's whole-volume scoring (one count launch + one loss launch) must equal the reference's per-(slice, object)
    accumulation, restated here with the numpy metric oracle on the driver's own predictions."""
    from medsam2_b200.validation import THRESHOLD, validation_sam
    from oracle.eval_seg import bce_with_logits_np, eval_seg_np
    size, T = 512, 3
    m = _build("sam2_hiera_t", video=True, image_size=size)
    packs = []
    for seed in (1234, 99):
        vol, boxes = btcv_volume(T, size, seed, 1)
        g = torch.Generator().manual_seed(seed)
        label = {f: {1: (torch.rand(1, size, size, generator=g) > 0.5).float()} for f in range(T)}
        del label[1][1]                                                     # no annotation on slice 1 -> all background
        packs.append({"image": vol[None], "label": label, "bbox": {f: {1: torch.tensor(boxes[f][0])} for f in range(T)}})
    packs.append({"image": btcv_volume(T, size, 5, 1)[0], "label": {f: {} for f in range(T)}, "bbox": {}})
    seen = []
    orig = m.propagate_in_video

    def spy(state, start_frame_idx=0):
        seen.append({})
        for f, ids, logits in orig(state, start_frame_idx=start_frame_idx):
            seen[-1][f] = logits.clone()
            yield f, ids, logits
    m.propagate_in_video = spy
    loss, (iou, dice) = validation_sam(m, packs, prompt="bbox", prompt_freq=2, device="cpu", fused_scoring=False)
    # default: scored straight from the tracker's low-res logits (up-sampling fused into the scoring pass, ms2_score_lowres):
    # same integers -> identical metrics; the loss differs by fp32 summation order only
    m.propagate_in_video = orig
    loss_f, (iou_f, dice_f) = validation_sam(m, packs, prompt="bbox", prompt_freq=2, device="cpu")
    assert iou_f == iou and dice_f == dice, (iou_f, iou, dice_f, dice)
    assert abs(loss_f - loss) <= 2e-6 * abs(loss), (loss_f, loss)
    want = np.zeros(3)
    for v in range(2):
        acc = np.zeros(3)
        for f in range(T):
            pred = seen[v][f][0][None].numpy()
            gt = packs[v]["label"][f].get(1, torch.zeros(1, size, size))[None].numpy()
            i_, d_ = eval_seg_np(pred, gt, THRESHOLD)
            acc += [bce_with_logits_np(pred, gt, 2.0), i_, d_]
        want += acc / T
    want /= 3                                                               # n_val counts the skipped volume
    assert abs(iou - want[1]) < 1e-12 and abs(dice - want[2]) < 1e-9, (iou, dice, want)
    assert abs(loss - want[0]) < 1e-6 * abs(want[0]), (loss, want[0])
    with pytest.raises(ValueError):
        validation_sam(m, packs[:1], prompt="scribble", device="cpu")


def test_product_refuses_cpu_without_backend():
    """No fake backend installed -> the hot path must fail loudly on CPU tensors (no fallback)."""
    from medsam2_b200 import ops
    from medsam2_b200.native import NativeError
    with pytest.raises(NativeError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-6)


def test_slice_encode_prefetch_planner():
    """Request-stride learning of the slice-encoding prefetcher (host logic only: the encoder is stubbed).  Config-3
    call order on 96 slices, batches of 8: prompts on the even slices, then tracking of the odd ones.  Every slice is
    encoded exactly once, in full batches, the odd slices are only encoded during tracking, and each batch is
    requested at least one request before its first slice is needed (after the very first one)."""
    from medsam2_b200.sam2_video_predictor import SAM2VideoPredictor

    class Stub:
        feature_cache_size, feature_encode_batch, feature_prefetch = 96, 8, True
        _prefetch_enabled = lambda self, st: True
        _plan_frames = SAM2VideoPredictor._plan_frames
        _get_image_feature = SAM2VideoPredictor._get_image_feature

        def __init__(self):
            self.batches = []

        def _encode_frames(self, st, frames, side):
            assert side and frames and all(f not in st["cached_features"] for f in frames)
            self.batches.append((st["_last_feature_request"], tuple(frames)))
            for f in frames:
                st["cached_features"][f] = (torch.zeros(1, 3, 4, 4), {"backbone_fpn": [torch.zeros(1, 2, 4, 4)],
                                                                     "vision_pos_enc": [torch.zeros(1, 2, 4, 4)]}, None)

        def _prepare_backbone_features(self, d):
            return d, None, None, None

    p = Stub()
    st = {"cached_features": {}, "num_frames": 96, "device": torch.device("cpu")}
    p._get_image_feature(st, 0, 1)                       # init_state warm-up
    for f in range(0, 96, 2):                            # add_new_bbox on every other slice
        p._get_image_feature(st, f, 1)
    n_phase1 = len(p.batches)
    assert all(f % 2 == 0 or f < 16 for _, fr in p.batches for f in fr), p.batches
    for f in range(0, 96, 2):                            # preflight: memory encoder on the prompted slices
        p._get_image_feature(st, f, 1)
    assert len(p.batches) == n_phase1
    for f in range(1, 96, 2):                            # tracking: only the un-prompted slices ask for features
        p._get_image_feature(st, f, 1)
    enc = [f for _, fr in p.batches for f in fr]
    assert sorted(enc) == list(range(96)) and len(p.batches) == 12 and all(len(fr) == 8 for _, fr in p.batches)
    for at, fr in p.batches[1:]:
        assert at != fr[0], (at, fr)                     # launched ahead of the request that needs it
    # reverse tracking plans downwards
    p2 = Stub()
    st2 = {"cached_features": {}, "num_frames": 20, "device": torch.device("cpu"), "prefetch_reverse": True}
    p2._get_image_feature(st2, 19, 1)
    assert p2.batches[0][1] == tuple(range(19, 11, -1)), p2.batches


def test_jpeg_ingest_host_path(tmp_path):
    """`load_video_frames` / `AsyncVideoFrameLoader` host side (thread-pool decode into the staging buffer, chunk
    bookkeeping, out-of-order requests, error propagation) with frames kept on the host (`offload_video_to_cpu`),
    against the real reference's frames (tests/golden/ingest_jpeg.npz)."""
    from medsam2_b200.utils.misc import AsyncVideoFrameLoader, load_video_frames
    z = np.load(f"{G}/ingest_jpeg.npz")
    for i in range(5):
        (tmp_path / f"{i}.jpg").write_bytes(z[f"jpeg_{i}"].tobytes())
    (tmp_path / "notes.txt").write_text("ignored")
    frames, h, w = load_video_frames(str(tmp_path), image_size=64, offload_video_to_cpu=True, device="cpu")
    assert (h, w) == tuple(z["hw"])
    _close(frames, z["sync"], 1e-6, "sync")
    lazy, h, w = load_video_frames(str(tmp_path), image_size=64, offload_video_to_cpu=True, async_loading_frames=True, device="cpu")
    assert isinstance(lazy, AsyncVideoFrameLoader) and len(lazy) == 5
    for i in (4, 0, 2, 3, 1):
        _close(lazy[i], z["async"][i], 1e-6, f"async {i}")
    lazy2 = AsyncVideoFrameLoader([str(tmp_path / f"{i}.jpg") for i in range(5)], 64, True, device="cpu", chunk=2, workers=3)
    _close(torch.stack([lazy2[i] for i in range(5)]), z["async"], 1e-6, "chunk 2")
    (tmp_path / "7.jpg").write_bytes(b"broken")
    lazy3, _, _ = load_video_frames(str(tmp_path), image_size=64, offload_video_to_cpu=True, async_loading_frames=True, device="cpu")
    with pytest.raises(RuntimeError, match="frame loading"):
        lazy3[5]
    with pytest.raises(RuntimeError, match="no images"):
        d2 = tmp_path / "empty"
        d2.mkdir()
        load_video_frames(str(d2), image_size=64)


def test_dense_layout_predicate_and_prompt_staging_fallback():
    """ops._dense decides whether a graph input / result can move as a raw byte copy (ms2_multi_copy); a host-to-host
    `to_device_async` is the plain `.to` (the pinned staging ring only exists for CUDA destinations)."""
    import torch
    from medsam2_b200 import ops
    from medsam2_b200.utils.misc import to_device_async
    assert ops._dense(torch.zeros(4, 5, 6))
    assert ops._dense(torch.zeros(2, 8, 16, 16).permute(0, 3, 1, 2))          # channels-last view: dense, not contiguous
    assert ops._dense(torch.zeros(4, 1, 5).transpose(0, 2))
    assert ops._dense(torch.zeros(0))
    assert not ops._dense(torch.zeros(16, 16)[:, ::2])                       # gaps
    assert not ops._dense(torch.zeros(3, 1).expand(3, 4))                    # overlapping (stride 0)
    assert not ops._dense(torch.zeros(8, 8)[:4, :4])                         # window of a larger buffer
    t = torch.arange(6, dtype=torch.float32).view(1, 3, 2)
    out = to_device_async(t, "cpu")
    assert out.device.type == "cpu" and torch.equal(out, t)
