"""GPU: parity of the video path AT THE BENCHMARKED GEOMETRY (BASELINE configs[2]: hiera_s, 1024², bbox prompt
every 2 slices, num_maskmem=7, fill_hole_area=8) against the oracle on the same device — teacher-forced per frame
and free-running (tests/parity_utils.py), in fp32 mode and in bf16 mode.

Tolerances are the north star's: mask logits (low-res AND video-res) <= 1e-4 abs in fp32 mode (teacher-forced),
<= 1e-2 abs in bf16 mode (teacher-forced and free-running).  Measured values are recorded by conftest
(`MS2_PARITY_TABLE`) and committed as profiles/r2_parity_table.txt."""
import pytest
import torch

import parity_utils as pu
from conftest import record_parity

pytestmark = pytest.mark.gpu

T, EVERY, SIZE, CFG = 32, 2, 1024, "sam2_hiera_s"


@pytest.fixture(scope="module")
def oracle32():
    vp, st, vid, vol, boxes = pu.oracle_run(CFG, SIZE, T, EVERY)
    return vp, st, vid, vol, boxes


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_teacher_forced_1024_32slices(oracle32, dt):
    import medsam2_b200
    vp, ost, ovid, vol, boxes = oracle32
    with medsam2_b200.compute(dt):
        m = pu.build_product(CFG, SIZE, T)
        r = pu.teacher_forced(m, vp, ost, vol, boxes, SIZE, T, EVERY)
    tag = f"config3_1024_{T}slices_teacher_forced_{'fp32' if dt == torch.float32 else 'bf16'}"
    record_parity(tag, {k: v for k, v in r.items() if k != "per_frame_low_res"})
    tol = 1e-4 if dt == torch.float32 else 1e-2
    assert r["cond_low_res"] <= tol, r
    assert r["tracked_low_res"] <= tol, r
    assert r["tracked_video_res"] <= tol, r
    # memory features / pointers feed the next frames: fp32 1e-3 abs on O(1) features; bf16 8e-2 (bf16 GEMM operands on
    # features of magnitude ~4: half an ulp is 1.6e-2)
    ftol = 1e-3 if dt == torch.float32 else 8e-2
    assert r["cond_maskmem"] <= ftol and r["tracked_maskmem"] <= ftol, r
    assert r["cond_obj_ptr"] <= ftol and r["tracked_obj_ptr"] <= ftol, r


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_free_running_1024_32slices(oracle32, dt):
    import medsam2_b200
    vp, ost, ovid, vol, boxes = oracle32
    with medsam2_b200.compute(dt):
        m = pu.build_product(CFG, SIZE, T)
        pst, pvid = pu.product_free_run(m, vol, boxes, SIZE, T, EVERY)
    w = pu.compare_free_running(ost, ovid, pst, pvid, T)
    record_parity(f"config3_1024_{T}slices_free_running_{'fp32' if dt == torch.float32 else 'bf16'}", w)
    tol = 1e-3 if dt == torch.float32 else 1e-2           # fp32 free-running: 16 frames of feedback on top of 1e-4
    assert w["low_res"] <= tol and w["video_res"] <= tol, w
    assert w["flipped_abs_logit_max"] <= tol, w            # masks differ only where the oracle's logit is within tol of 0
    assert w["sign_agree_min"] >= (0.9999 if dt == torch.float32 else 0.99), w


def test_free_running_bf16_vs_bf16_autocast_oracle():
    """the reference's own precision (train_3d.py:28,57: bf16 autocast) as the checker: both sides carry bf16 rounding,
    so this bounds the distance between two bf16 implementations, not the distance to the exact answer."""
    import medsam2_b200
    vp, ost, ovid, vol, boxes = pu.oracle_run(CFG, SIZE, 16, EVERY, autocast=True)
    with medsam2_b200.compute(torch.bfloat16):
        m = pu.build_product(CFG, SIZE, 16)
        pst, pvid = pu.product_free_run(m, vol, boxes, SIZE, 16, EVERY)
    w = pu.compare_free_running(ost, ovid, pst, pvid, 16)
    record_parity("config3_1024_16slices_free_running_bf16_vs_autocast_oracle", w)
    assert w["low_res"] <= 2e-2 and w["flipped_abs_logit_max"] <= 2e-2 and w["sign_agree_min"] >= 0.98, w
