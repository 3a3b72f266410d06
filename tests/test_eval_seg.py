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

from oracle.eval_seg import eval_seg_np

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
