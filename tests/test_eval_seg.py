"""Segmentation metrics (`func_3d/utils.py:139-240` eval_seg / iou / dice_coeff).

CPU: the numpy oracle (oracle/eval_seg.py) and the product's host arithmetic (medsam2_b200/utils/eval.py `_reduce`)
against the answers of the REAL reference (tests/golden/eval_seg_cases.npz, made by make_golden_eval.py).
GPU: the one-pass count kernel `ms2_seg_counts` bit-exact against numpy counts (ragged, unaligned, empty, full-size
planes) and `eval_seg` / `eval_seg_frames` equal to the golden answers.  The bar is exact equality: the counts are
integers and the host arithmetic repeats the reference's float64 / fp32 operations in the same order."""
import os

import numpy as np
import pytest
import torch

from oracle.eval_seg import bce_with_logits_np, eval_seg_np

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["c1_b1", "c1_b3", "c2_b2", "c1_odd", "c1_empty", "c1_full", "c3_b2_one_thr", "c1_hard01"]


def _case(name):
    z = np.load(f"{G}/eval_seg_cases.npz")
    return z[name + "/pred"], z[name + "/gt"], tuple(float(t) for t in z[name + "/thr"]), z[name + "/res"]


def _counts_np(pred, gt, thr):
    """int64 [b,c,T,3] by plain numpy (the integer statement of the kernel)."""
    out = np.zeros(pred.shape[:2] + (len(thr), 3), np.int64)
    for t, th in enumerate(thr):
        a, b = pred > np.float32(th), gt > np.float32(th)
        out[:, :, t, 0] = (a & b).sum((2, 3))
        out[:, :, t, 1] = a.sum((2, 3))
        out[:, :, t, 2] = b.sum((2, 3))
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_answers(name):
    pred, gt, thr, res = _case(name)
    got = np.array([float(r) for r in eval_seg_np(pred, gt, thr)])
    assert np.array_equal(got, res), (name, got, res)


@pytest.mark.parametrize("name", CASES)
def test_host_reduction_matches_reference_answers(name):
    """the product's count -> (IoU, Dice) arithmetic, fed numpy counts (no GPU involved)"""
    from medsam2_b200.utils.eval import _reduce
    pred, gt, thr, res = _case(name)
    got = np.array([float(r) for r in _reduce(_counts_np(pred, gt, thr))])
    assert np.array_equal(got, res), (name, got, res)


@pytest.mark.parametrize("name", CASES)
def test_oracle_bce_matches_reference_loss(name):
    """float64 restatement vs `BCEWithLogitsLoss(pos_weight=2)` of the reference (CPU fp32): relative 2e-6"""
    z = np.load(f"{G}/eval_seg_cases.npz")
    pred, gt, want = z[name + "/pred"], z[name + "/gt"], z[name + "/bce"]
    got = [bce_with_logits_np(pred[i:i + 1], gt[i:i + 1]) for i in range(len(pred))] + [bce_with_logits_np(pred, gt)]
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-9)


def test_product_refuses_cpu_tensors():
    from medsam2_b200 import native
    from medsam2_b200.utils.eval import eval_seg
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(native.NativeError):
        eval_seg(x, x, (0.5,))
    with pytest.raises(ValueError):
        eval_seg(x, torch.zeros(1, 1, 8, 4), (0.5,))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_eval_seg_equals_reference_answers(name):
    from medsam2_b200.utils.eval import eval_seg
    pred, gt, thr, res = _case(name)
    got = np.array([float(r) for r in eval_seg(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), thr)])
    assert np.array_equal(got, res), (name, got, res)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (3, 1, 7, 5), (2, 2, 37, 53), (5, 1, 64, 64), (2, 1, 1024, 1024),
                                   (96, 1, 256, 256), (1, 1, 3, 4099)])
def test_gpu_counts_bit_exact(shape):
    from medsam2_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    pred = torch.randn(shape, generator=g)
    gt = torch.rand(shape, generator=g)
    pred[0, 0, 0, 0] = float("nan")                       # NaN compares false on both sides
    thr = (0.1, 0.3, 0.5, 0.7, 0.9)
    b, c = shape[:2]
    got = ops.seg_counts(pred.cuda().reshape(b * c, -1), gt.cuda().reshape(b * c, -1), thr).cpu().numpy()
    assert np.array_equal(got.reshape(b, c, 5, 3), _counts_np(pred.numpy(), gt.numpy(), thr))


@pytest.mark.gpu
def test_gpu_counts_unaligned_and_limits():
    from medsam2_b200 import native, ops
    g = torch.Generator().manual_seed(7)
    base_p, base_g = torch.randn(4 * 1000 + 1, generator=g).cuda(), torch.rand(4 * 1000 + 1, generator=g).cuda()
    p, q = base_p[1:].view(4, 1000), base_g[1:].view(4, 1000)          # rows of 1000 floats starting 4 bytes off
    thr = tuple(np.linspace(-1, 1, 8))
    got = ops.seg_counts(p, q, thr).cpu().numpy()
    want = _counts_np(p.cpu().numpy()[:, None, None], q.cpu().numpy()[:, None, None], thr)[:, 0]
    assert np.array_equal(got, want)
    assert ops.seg_counts(p[:0], q[:0], thr).shape == (0, 8, 3)
    z = ops.seg_counts(p[:, :0], q[:, :0], (0.5,))
    assert z.shape == (4, 1, 3) and int(z.abs().sum()) == 0
    with pytest.raises(ValueError):
        ops.seg_counts(p, q, tuple(range(9)))
    with pytest.raises(native.NativeError):
        ops.seg_counts(p.t(), q.t(), (0.5,))                            # non-contiguous


@pytest.mark.gpu
def test_gpu_eval_seg_frames_and_helpers():
    """a whole volume in one launch == the per-slice calls of func_3d/function.py:276-305; more than 8 thresholds
    split over launches; iou / dice_coeff helpers agree with the oracle"""
    from medsam2_b200.utils.eval import dice_coeff, eval_seg, eval_seg_frames, iou
    from oracle.eval_seg import dice_coeff_np, iou_np
    g = torch.Generator().manual_seed(11)
    gt = (torch.rand(6, 1, 96, 96, generator=g) > 0.6).float()
    pred = gt * 3 - 1 + torch.randn(6, 1, 96, 96, generator=g)
    thr = (0.1, 0.3, 0.5, 0.7, 0.9)
    per_frame = eval_seg_frames(pred.cuda(), gt.cuda(), thr)
    for i in range(6):
        assert per_frame[i] == eval_seg(pred[i:i + 1].cuda(), gt[i:i + 1].cuda(), thr)
        assert per_frame[i] == eval_seg_np(pred[i:i + 1].numpy(), gt[i:i + 1].numpy(), thr)
    thr11 = tuple(np.linspace(-2, 2, 11))
    assert eval_seg(pred.cuda(), gt.cuda(), thr11) == eval_seg_np(pred.numpy(), gt.numpy(), thr11)
    a, b = (pred[:, 0] > 0).float(), gt[:, 0]
    assert iou(a.cuda(), b.cuda()) == iou_np(a.numpy().astype("int32"), b.numpy().astype("int32"))
    d = dice_coeff(a.cuda(), b.cuda())
    assert d.is_cuda and d.shape == (1,) and float(d) == dice_coeff_np(a.numpy(), b.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_bce_matches_reference_loss(name):
    """validation loss kernel vs the golden `BCEWithLogitsLoss(pos_weight=2)` values and the float64 oracle.
    Tolerance: relative 2e-6 against the fp32 reference values (its own summation error), 5e-7 against the oracle."""
    from medsam2_b200.utils.eval import bce_with_logits, bce_with_logits_frames
    z = np.load(f"{G}/eval_seg_cases.npz")
    pred, gt, want = z[name + "/pred"], z[name + "/gt"], z[name + "/bce"]
    p, g = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    per = bce_with_logits_frames(p, g, 2.0)
    assert per.is_cuda and per.dtype == torch.float32 and per.shape == (len(pred),)
    whole = bce_with_logits(p, g, 2.0)
    got = np.concatenate([per.cpu().numpy().astype(np.float64), [float(whole)]])
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-9)
    oracle = [bce_with_logits_np(pred[i:i + 1], gt[i:i + 1]) for i in range(len(pred))] + [bce_with_logits_np(pred, gt)]
    np.testing.assert_allclose(got, oracle, rtol=5e-7, atol=1e-9)


@pytest.mark.gpu
def test_gpu_bce_edges():
    """saturated logits stay finite, pos_weight other than 2, ragged / unaligned rows, empty inputs, full-size planes"""
    from medsam2_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.tensor([[-100.0, 100.0, 0.0, 88.0, -88.0, 1e4, -1e4]]).cuda()
    y = torch.tensor([[1.0, 0.0, 0.5, 1.0, 0.0, 1.0, 0.0]]).cuda()
    for pw in (1.0, 2.0, 0.25):
        got = float(ops.bce_logits_sum(x, y, pw)[0])
        want = bce_with_logits_np(x.cpu().numpy(), y.cpu().numpy(), pw) * 7
        assert np.isfinite(got) and abs(got - want) <= 1e-6 * abs(want), (pw, got, want)
        ref = torch.nn.functional.binary_cross_entropy_with_logits(x.cpu(), y.cpu(), pos_weight=torch.tensor([pw]), reduction="sum")
        assert abs(got - float(ref)) <= 2e-6 * abs(float(ref))
    base_x, base_y = torch.randn(3 * 1001 + 1, generator=g).cuda(), torch.rand(3 * 1001 + 1, generator=g).cuda()
    xs, ys = base_x[1:].view(3, 1001), base_y[1:].view(3, 1001)
    got = ops.bce_logits_sum(xs, ys, 2.0).cpu().numpy()
    want = [bce_with_logits_np(xs[i].cpu().numpy(), ys[i].cpu().numpy()) * 1001 for i in range(3)]
    np.testing.assert_allclose(got, want, rtol=5e-7)
    assert ops.bce_logits_sum(xs[:0], ys[:0], 2.0).shape == (0,)
    assert float(ops.bce_logits_sum(xs[:, :0], ys[:, :0], 2.0).abs().sum()) == 0.0
    big_y = (torch.rand(4, 1024 * 1024, generator=g) > 0.7).float()
    big_x = (big_y * 4 - 1.5 + torch.randn(4, 1024 * 1024, generator=g))
    got = ops.bce_logits_sum(big_x.cuda(), big_y.cuda(), 2.0).cpu().numpy() / (1024 * 1024)
    want = [bce_with_logits_np(big_x[i].numpy(), big_y[i].numpy()) for i in range(4)]
    np.testing.assert_allclose(got, want, rtol=5e-7)
