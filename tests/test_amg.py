"""Automatic mask generator (reference sam2_train/automatic_mask_generator.py + utils/amg.py).

Golden answers come from the REAL reference (tests/golden/make_golden.py amg -> amg_hiera_t_1024.npz): `generate()` records
of a seeded hiera_t model on a 300x400 image (with crops, without NMS, with mask-to-mask refinement) and known answers of
the helper functions (stability score, boxes, RLE, torchvision NMS, crop boxes, point grids).
CPU: helpers and the host logic of `generate()` with the native ops replaced by their torch statements (tests/ref_ops.py).
GPU: the two kernels bit-exact against those statements, the drop-in helpers equal to the golden answers, `generate()`
in fp32 mode against the golden records.  Discrete outputs (areas, boxes) are compared with the tolerances written below:
fp32 rounding moves a few boundary pixels of the low-amplitude random-weight logits."""
import os

import numpy as np
import pytest
import torch

import ref_ops
from oracle.config import get_config
from oracle.weights import make_state_dict

G = os.path.join(os.path.dirname(__file__), "golden")
Z = lambda: np.load(f"{G}/amg_hiera_t_1024.npz")
AMG_KW = dict(points_per_side=4, points_per_batch=8, pred_iou_thresh=0.5216, stability_score_thresh=0.65,
              stability_score_offset=0.02, mask_threshold=0.0, box_nms_thresh=0.95, crop_n_layers=1, crop_nms_thresh=0.95,
              crop_n_points_downscale_factor=2, output_mode="uncompressed_rle")
VARIANTS = {"plain": {}, "nonms": dict(box_nms_thresh=1.0, crop_nms_thresh=1.0),
            "m2m": dict(use_m2m=True, crop_n_layers=0, box_nms_thresh=1.0, pred_iou_thresh=0.0, stability_score_thresh=0.0)}


def amg_image(h=300, w=400, seed=3):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for _ in range(6):
        cy, cx, r = g.uniform(0, h), g.uniform(0, w), g.uniform(30, 90)
        col = g.uniform(40, 255, size=3)
        img += np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))[..., None] * col
    return np.clip(img, 0, 255).astype(np.uint8)


def _build(device):
    import medsam2_b200
    m = medsam2_b200.build_sam2("sam2_hiera_t", device=device)
    m.load_state_dict(make_state_dict(get_config("sam2_hiera_t")), strict=True)
    return m


def _check_records(anns, z, tag, area_rtol, box_atol, iou_atol, stab_atol):
    """same candidates as the reference (matched by prompt point, then by closeness of area / IoU / stability: predicted IoUs of different prompts can
    be closer than the tolerance, so the ORDER of near-ties is not compared), each within the stated tolerances"""
    n = int(z[f"{tag}/n"])
    assert len(anns) == n, (tag, len(anns), n)
    left = list(range(n))
    for a in anns:
        assert set(a) == {"segmentation", "area", "bbox", "predicted_iou", "point_coords", "stability_score", "crop_box"}
        cand = [i for i in left if np.allclose(a["point_coords"][0], z[f"{tag}/point_coords"][i])]
        assert cand, (tag, a["point_coords"], "no reference record left for this prompt")
        i = min(cand, key=lambda j: abs(a["area"] - z[f"{tag}/area"][j]) / z[f"{tag}/area"][j]
                + abs(a["predicted_iou"] - z[f"{tag}/predicted_iou"][j]) + abs(a["stability_score"] - z[f"{tag}/stability_score"][j]))
        left.remove(i)
        assert a["crop_box"] == z[f"{tag}/crop_box"][i].tolist()
        assert abs(a["area"] - z[f"{tag}/area"][i]) <= area_rtol * z[f"{tag}/area"][i], (tag, i, a["area"], z[f"{tag}/area"][i])
        assert np.abs(np.array(a["bbox"]) - z[f"{tag}/bbox"][i]).max() <= box_atol, (tag, i, a["bbox"], z[f"{tag}/bbox"][i])
        assert abs(a["predicted_iou"] - z[f"{tag}/predicted_iou"][i]) <= iou_atol, (tag, i)
        assert abs(a["stability_score"] - z[f"{tag}/stability_score"][i]) <= stab_atol, (tag, i)
        rle = a["segmentation"]
        assert rle["size"] == [300, 400] and sum(rle["counts"]) == 300 * 400 and sum(rle["counts"][1::2]) == a["area"]
    ious = [a["predicted_iou"] for a in anns]
    assert all(ious[k] >= ious[k + 1] - 1e-6 for k in range(len(ious) - 1)) or tag == "plain", "records come out by descending predicted IoU"


# ------------------------------------------------------------------------------------------------ CPU
def test_host_helpers_match_reference():
    from medsam2_b200.utils import amg
    z = Z()
    assert np.array_equal(amg.build_point_grid(3), z["h/grid3"])
    cb, li = amg.generate_crop_boxes((300, 400), 2, 512 / 1500)
    assert np.array_equal(np.array(cb), z["h/crop_boxes"]) and np.array_equal(np.array(li), z["h/crop_layers"])
    for thr in (0.3, 0.7):
        assert np.array_equal(amg.nms(z["h/nms_boxes"], z["h/nms_scores"], thr), z[f"h/nms_keep_{thr}"])
    near = amg.is_box_near_crop_edge(np.array([[0, 0, 50, 50], [30, 5, 99, 60], [25, 25, 60, 60], [0, 40, 60, 99]]),
                                     [100, 0, 200, 100], [0, 0, 400, 100])
    assert np.array_equal(near, z["h/near_edge"])
    logits = torch.from_numpy(z["h/logits"])
    st = ref_ops.mask_stats(logits, 0.1, 0.5).numpy().astype(np.int64)
    assert np.array_equal(amg.stability_from_stats(st), z["h/stability"], equal_nan=True)
    assert np.array_equal(amg.boxes_from_stats(st), z["h/boxes"])
    mt = ref_ops.mask_binarize_t(logits, torch.arange(5, dtype=torch.int32), 0.1, (37, 53), (0, 0)).numpy()
    rles = [amg.rle_from_transposed(mt[i]) for i in range(5)]
    assert [len(r["counts"]) for r in rles] == z["h/rle_lens"].tolist()
    assert np.array_equal(np.concatenate([r["counts"] for r in rles]), z["h/rle_counts"])
    assert [amg.area_from_rle(r) for r in rles] == z["h/areas"].tolist()
    for i, r in enumerate(rles):
        assert np.array_equal(amg.rle_to_mask(r), (logits[i] > 0.1).numpy())
    assert amg.box_xyxy_to_xywh(np.array([3, 4, 10, 20])).tolist() == [3, 4, 7, 16]


def test_remove_small_regions_matches_reference():
    """scipy 8-connected labelling in place of OpenCV's: identical masks and `changed` flags (amg_regions.npz holds the
    answers of the reference's `remove_small_regions` on seeded disc + salt-and-pepper masks, empty, full, tiny islands)"""
    from medsam2_b200.utils import amg
    z = np.load(f"{G}/amg_regions.npz")
    for i in range(int(z["n"])):
        for mode in ("holes", "islands"):
            for thr in (5, 40):
                out, changed = amg.remove_small_regions(z[f"mask_{i}"], thr, mode)
                assert bool(changed) == bool(z[f"changed_{i}_{mode}_{thr}"]), (i, mode, thr)
                assert np.array_equal(np.asarray(out, bool), z[f"out_{i}_{mode}_{thr}"]), (i, mode, thr)


def test_postprocess_small_regions_host():
    """the optional clean-up pass: islands below the area threshold disappear, the box is recomputed, a duplicate that
    needed no change is preferred by the NMS"""
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator, _Candidates
    from medsam2_b200.utils import amg
    big = np.zeros((64, 80), bool); big[10:40, 20:60] = True
    noisy = big.copy(); noisy[50, 70] = True; noisy[20, 30] = False          # one island pixel, one hole pixel
    other = np.zeros((64, 80), bool); other[45:60, 5:15] = True
    data = _Candidates()
    data.rles = [amg.rle_from_transposed(np.ascontiguousarray(m.T).astype(np.uint8)) for m in (noisy, big, other)]
    data.boxes = np.array([[20, 10, 70, 50], [20, 10, 59, 39], [5, 45, 14, 59]], np.int64)
    data.iou_preds = np.array([0.9, 0.8, 0.7], np.float32)
    data.points = np.zeros((3, 2)); data.stability_score = np.ones(3, np.float32); data.crop_boxes = np.zeros((3, 4), np.int64)
    out = SAM2AutomaticMaskGenerator.postprocess_small_regions(data, min_area=4, nms_thresh=0.7)
    assert len(out) == 2                                                     # the cleaned noisy mask duplicates `big`
    kept = [amg.rle_to_mask(r) for r in out.rles]
    assert any(np.array_equal(k, big) for k in kept) and any(np.array_equal(k, other) for k in kept)
    assert sorted(out.iou_preds.tolist()) == [np.float32(0.7), np.float32(0.8)]        # the unchanged duplicate won


def test_device_rle_host_side(monkeypatch):
    """`rles_from_device` (positions compacted by `ops.rle_transitions`, here its torch statement) == the plain host
    encoder, including the re-run when a mask has more transitions than the first capacity"""
    from medsam2_b200.utils import amg
    ref_ops.install(monkeypatch)
    z = Z()
    logits = torch.from_numpy(z["h/logits"])
    mt = ref_ops.mask_binarize_t(logits, torch.arange(5, dtype=torch.int32), 0.1, (37, 53), (0, 0))
    for cap in (4096, 7):
        rles = amg.rles_from_device(mt, cap=cap)
        assert rles == [amg.rle_from_transposed(mt[i].numpy()) for i in range(5)]
        assert np.array_equal(np.concatenate([r["counts"] for r in rles]), z["h/rle_counts"])
    assert amg.rles_from_device(mt[:0]) == []


@pytest.mark.parametrize("tag", ["plain", "m2m"])
def test_generate_host_logic(monkeypatch, tag):
    """`generate()` with the native ops replaced by their torch statements vs the records of the real reference."""
    import medsam2_b200  # noqa: F401
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
    ref_ops.install(monkeypatch)
    gen = SAM2AutomaticMaskGenerator(_build("cpu"), **{**AMG_KW, **VARIANTS[tag]})
    _check_records(gen.generate(amg_image()), Z(), tag, area_rtol=1.5e-2, box_atol=1, iou_atol=5e-4, stab_atol=5e-3)


def test_generate_without_survivors(monkeypatch):
    """nothing passes the predicted-IoU filter -> an empty list, in every output mode that needs no extra package"""
    import medsam2_b200  # noqa: F401
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
    ref_ops.install(monkeypatch)
    m = _build("cpu")
    for mode in ("binary_mask", "uncompressed_rle"):
        gen = SAM2AutomaticMaskGenerator(m, points_per_side=2, points_per_batch=4, pred_iou_thresh=0.99, output_mode=mode)
        assert gen.generate(amg_image(64, 96)) == []


def test_constructor_contract():
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
    m = _build("cpu")
    with pytest.raises(AssertionError):
        SAM2AutomaticMaskGenerator(m, points_per_side=None, point_grids=None)
    with pytest.raises(AssertionError):
        SAM2AutomaticMaskGenerator(m, output_mode="png")
    with pytest.raises(ImportError):
        SAM2AutomaticMaskGenerator(m, output_mode="coco_rle")          # pycocotools is absent here, as for the reference
    g = SAM2AutomaticMaskGenerator(m, points_per_side=None, point_grids=[np.array([[0.5, 0.5]])])
    assert len(g.point_grids) == 1


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 37, 53), (3, 300, 400), (48, 256, 256), (2, 1024, 1024), (200, 8, 8),
                                   (3, 8, 2048)])
def test_gpu_mask_stats_bit_exact(shape):
    from medsam2_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    if shape[0] > 2:
        x[1] = -5.0                                        # empty plane
        x[2] = 5.0                                         # full plane
    for thr, off in ((0.0, 1.0), (0.1, 0.5), (2.5, 0.0)):
        got = ops.mask_stats(x.cuda(), thr, off).cpu()
        assert torch.equal(got, ref_ops.mask_stats(x, thr, off)), (shape, thr, off)
    assert ops.mask_stats(x.cuda()[:0], 0.0, 1.0).shape == (0, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,canvas,origin", [((5, 37, 53), (37, 53), (0, 0)), ((4, 100, 130), (300, 400), (270, 17)),
                                                 ((3, 300, 400), (300, 400), (0, 0)), ((2, 33, 31), (64, 64), (33, 31)),
                                                 ((2, 1024, 1024), (1024, 1024), (0, 0))])
def test_gpu_mask_binarize_t_bit_exact(shape, canvas, origin):
    from medsam2_b200 import native, ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    sel = torch.tensor([shape[0] - 1, 0], dtype=torch.int32)
    got = ops.mask_binarize_t(x.cuda(), sel.cuda(), 0.1, canvas, origin).cpu()
    assert torch.equal(got, ref_ops.mask_binarize_t(x, sel, 0.1, canvas, origin))
    assert ops.mask_binarize_t(x.cuda(), sel.cuda()[:0], 0.1, canvas, origin).shape == (0, canvas[1], canvas[0])
    with pytest.raises(native.NativeError):
        ops.mask_binarize_t(x.cuda(), sel.cuda(), 0.1, (shape[1] - 1, shape[2]), (0, 0))      # crop larger than canvas


@pytest.mark.gpu
def test_gpu_rle_transitions_bit_exact():
    """device run-length boundaries vs the host encoder: noise (capacity overflow -> re-run), blobs, constant masks,
    lengths that are not a multiple of 16, a base pointer that is not 16-byte aligned, more than one chunk per CTA"""
    from medsam2_b200 import ops
    from medsam2_b200.utils import amg
    g = torch.Generator().manual_seed(9)
    cases = [(torch.rand(3, 53, 37, generator=g) > 0.5), (torch.rand(2, 400, 300, generator=g) > 0.98),
             torch.zeros(2, 64, 64, dtype=torch.bool), torch.ones(1, 33, 31, dtype=torch.bool)]
    yy, xx = torch.meshgrid(torch.arange(1024), torch.arange(1024), indexing="ij")
    cases.append(torch.stack([((yy - 300) ** 2 + (xx - 500) ** 2 < 200 ** 2), ((yy > 100) & (xx % 97 < 40))]))
    for m in cases:
        mt = m.to(torch.uint8).contiguous()
        want = [amg.rle_from_transposed(mt[i].numpy()) for i in range(len(mt))]
        assert amg.rles_from_device(mt.cuda(), cap=64) == want
        assert amg.rles_from_device(mt.cuda()) == want
    base = (torch.rand(2 * 1000 + 1, generator=g) > 0.7).to(torch.uint8).cuda()
    un = base[1:].view(2, 1000)                                           # rows start 1 byte off a 16-byte boundary
    pos, cnt = ops.rle_transitions(un, 1000)
    for k in range(2):
        want = np.flatnonzero(np.diff(un[k].cpu().numpy().astype(np.int8)) != 0) + 1
        assert int(cnt[k]) == len(want) and np.array_equal(pos[k, :len(want)].cpu().numpy(), want)
    pos, cnt = ops.rle_transitions(un, 0)                                  # counting only
    assert pos.shape == (2, 0) and int(cnt.sum()) > 0


@pytest.mark.gpu
def test_gpu_drop_in_helpers_equal_reference():
    from medsam2_b200.utils import amg
    z = Z()
    logits = torch.from_numpy(z["h/logits"]).cuda()
    s = amg.calculate_stability_score(logits, 0.1, 0.5)
    assert s.is_cuda and np.array_equal(s.cpu().numpy(), z["h/stability"], equal_nan=True)
    b = amg.batched_mask_to_box(logits > 0.1)
    assert b.is_cuda and np.array_equal(b.cpu().numpy(), z["h/boxes"])
    rles = amg.mask_to_rle_pytorch(logits > 0.1)
    assert np.array_equal(np.concatenate([r["counts"] for r in rles]), z["h/rle_counts"])
    assert [r["size"] for r in rles] == [[37, 53]] * 5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "nonms", "m2m"])
def test_gpu_generate_fp32(tag):
    """CUDA path in fp32 mode vs the records of the real reference: same candidates in the same order; area within 3 %,
    box within 2 px, predicted IoU within 2e-3, stability within 1e-2."""
    import medsam2_b200
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
    with medsam2_b200.compute(torch.float32):
        gen = SAM2AutomaticMaskGenerator(_build("cuda"), **{**AMG_KW, **VARIANTS[tag]})
        anns = gen.generate(amg_image())
    _check_records(anns, Z(), tag, area_rtol=3e-2, box_atol=2, iou_atol=2e-3, stab_atol=1e-2)


@pytest.mark.gpu
def test_gpu_generate_bf16_records_are_consistent():
    """bf16 mode (the product default): thresholds of the golden run no longer separate the candidates, so only the
    internal consistency of every record is checked: RLE decodes to a mask with the recorded area and box."""
    from medsam2_b200.automatic_mask_generator import SAM2AutomaticMaskGenerator
    from medsam2_b200.utils import amg
    kw = {**AMG_KW, "pred_iou_thresh": 0.0, "stability_score_thresh": 0.3, "output_mode": "binary_mask"}
    anns = SAM2AutomaticMaskGenerator(_build("cuda"), **kw).generate(amg_image())
    assert len(anns) > 0
    for a in anns:
        m = a["segmentation"]
        assert m.shape == (300, 400) and m.dtype == bool and int(m.sum()) == a["area"]
        ys, xs = np.nonzero(m)
        assert a["bbox"] == [xs.min(), ys.min(), xs.max() - xs.min(), ys.max() - ys.min()]
        assert 0.3 <= a["stability_score"] <= 1.0
