"""GPU: `ms2_cc_label` (the C-ABI replacement of `sam2_train._C.get_connected_componnets`) against the REFERENCE's own
kernel compiled for sm_100a from /root/reference/sam2_train/csrc/connected_components.cu (oracle/build_ref.py ->
oracle/_ref/ref_cc/ref_cc.so, built in the container, shipped with the snapshot).  Labels and areas bit-exact at the
sizes the tracker and the image predictor use, plus the fill-holes decision derived from them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_cc():
    from oracle.build_ref import load_ref_cc
    mod = load_ref_cc()
    if mod is None:
        pytest.skip("oracle/_ref/ref_cc/ref_cc.so was not built (python oracle/build_ref.py in the build container)")
    return mod


@pytest.mark.parametrize("shape,density", [((13, 1, 256, 256), 0.5), ((1, 1, 1024, 1024), 0.55), ((2, 1, 1024, 1024), 0.93),
                                           ((3, 1, 256, 256), 0.08), ((4, 1, 64, 48), 0.6), ((1, 1, 512, 640), 0.4),
                                           ((2, 1, 256, 256), 0.0), ((2, 1, 256, 256), 1.0)])
def test_cc_label_matches_compiled_reference(ref_cc, shape, density):
    from medsam2_b200 import ops
    g = torch.Generator().manual_seed(int(density * 100) + shape[0])
    m = (torch.rand(shape, generator=g) < density).to(torch.uint8).cuda()
    rl, rc = ref_cc.get_connected_componnets(m)
    l, c = ops.cc_label(m)
    assert torch.equal(l, rl), "labels differ from the reference kernel"
    assert torch.equal(c, rc), "areas differ from the reference kernel"


def test_cc_structured_and_smooth_masks(ref_cc):
    """blobs / rings / diagonal chains at 1024^2 (what thresholded mask logits look like), through the public `_C` shim"""
    from medsam2_b200 import _C
    yy, xx = torch.meshgrid(torch.arange(1024), torch.arange(1024), indexing="ij")
    m = ((torch.sin(xx / 9.0) * torch.cos(yy / 7.0) + torch.sin((xx + yy) / 23.0)) > 0.3)
    m |= (xx == yy) | ((xx + yy) % 64 == 0)
    m = m.to(torch.uint8)[None, None].cuda()
    rl, rc = ref_cc.get_connected_componnets(m)
    l, c = _C.get_connected_componnets(m)
    assert torch.equal(l, rl) and torch.equal(c, rc)


def test_fill_holes_matches_reference_pipeline(ref_cc):
    """utils/misc.py:247-258 evaluated with the reference kernel's labels/areas == the fused ms2_fill_holes"""
    from medsam2_b200.utils.misc import fill_holes_in_mask_scores
    g = torch.Generator().manual_seed(3)
    base = torch.nn.functional.interpolate(torch.randn(5, 1, 32, 32, generator=g), size=(256, 256), mode="bilinear")
    x = (base + 0.8 * torch.randn(5, 1, 256, 256, generator=g) * (torch.rand(5, 1, 256, 256, generator=g) < 0.05)).cuda()
    labels, areas = ref_cc.get_connected_componnets((x <= 0).to(torch.uint8))
    want = torch.where((labels > 0) & (areas <= 8), 0.1, x)
    assert torch.equal(fill_holes_in_mask_scores(x, 8), want)
