"""GPU: every C-ABI kernel against its plain-torch statement (tests/ref_ops.py) on seeded inputs, and
connected components bit-exact against the oracle and the committed golden vectors.
Tolerances: fp32 paths 1e-4 abs unless stated (fp32 accumulate, different summation order);
bf16 operand paths are compared against the fp32 statement fed the SAME bf16-rounded operands."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ref_ops

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    from medsam2_b200 import ops as o
    return o


def gen(seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=gen(seed)) * scale).cuda()


def close(a, b, tol, what=""):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert err <= tol, f"{what}: max abs err {err} > {tol}"


def close_rel(a, b, atol, rtol, what=""):
    """|a-b| <= atol + rtol*|b| (bf16 outputs: half an ulp of a value near 4 is already 7.8e-3)."""
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    excess = ((a - b).abs() - rtol * b.abs()).max().item() if a.numel() else 0.0
    assert excess <= atol, f"{what}: max (abs err - {rtol}*|ref|) = {excess} > {atol}"


# ------------------------------------------------------------------ connected components (bit-exact)
def test_cc_golden_vectors(ops):
    z = np.load(f"{G}/cc_cases.npz")
    for i in range(int(z["n"])):
        m = torch.from_numpy(z[f"mask_{i}"])[None, None].cuda()
        l, c = ops.cc_label(m)
        assert np.array_equal(l[0, 0].cpu().numpy(), z[f"labels_{i}"]), i
        assert np.array_equal(c[0, 0].cpu().numpy(), z[f"counts_{i}"]), i


@pytest.mark.parametrize("shape,density", [((3, 1, 256, 256), 0.5), ((2, 1, 256, 256), 0.9), ((1, 1, 256, 256), 0.02),
                                           ((5, 1, 64, 48), 0.6), ((1, 1, 1024, 1024), 0.55), ((2, 1, 512, 640), 0.4),
                                           ((0, 1, 16, 16), 0.5), ((1, 1, 2, 2), 1.0)])
def test_cc_matches_oracle(ops, shape, density):
    from oracle.sam2_oracle import connected_components_np
    m = (torch.rand(shape, generator=gen(3)) < density).to(torch.uint8)
    l, c = ops.cc_label(m.cuda())
    lo, co = connected_components_np(m.numpy())
    assert np.array_equal(l.cpu().numpy(), lo)
    assert np.array_equal(c.cpu().numpy(), co)


def test_cc_properties_full_size(ops):
    """size-independent properties at 1024x1024: idempotent relabel, counts sum = #fg, labels constant
    on 8-neighbours."""
    m = (torch.rand((2, 1, 1024, 1024), generator=gen(9)) < 0.58).to(torch.uint8).cuda()
    l, c = ops.cc_label(m)
    fg = m.bool()
    assert (l[~fg] == 0).all() and (c[~fg] == 0).all() and (l[fg] > 0).all()
    for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
        a = l[..., : l.shape[-2] - dy, max(0, -dx): l.shape[-1] - max(0, dx)]
        b = l[..., dy:, max(0, dx): l.shape[-1] - max(0, -dx)]
        both = (a > 0) & (b > 0)
        assert (a[both] == b[both]).all()
    # every component's count equals its number of pixels
    for n in range(2):
        lab = l[n, 0][fg[n, 0]].long()
        cnt = c[n, 0][fg[n, 0]].long()
        hist = torch.bincount(lab)
        assert (hist[lab] == cnt).all()


def test_cc_entry_contract(ops):
    from medsam2_b200 import _C
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros(1, 1, 4, 4, dtype=torch.uint8))           # CPU tensor
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros(1, 1, 4, 4, dtype=torch.float32).cuda())  # dtype
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros(1, 1, 5, 4, dtype=torch.uint8).cuda())    # odd H
    out = _C.get_connected_componnets(torch.ones(1, 1, 4, 4, dtype=torch.uint8).cuda())
    assert out[0].dtype == torch.int32 and (out[0] == 1).all() and (out[1] == 16).all()


def test_fill_holes(ops):
    x = rnd(3, 1, 256, 256, seed=4)
    x = torch.where(torch.rand(x.shape, generator=gen(5)).cuda() < 0.03, -x.abs(), x.abs())   # sparse small holes
    close(ops.fill_holes(x, 8), ref_ops.fill_holes(x, 8), 0.0, "fill_holes")


@pytest.mark.parametrize("density,area", [(0.03, 8), (0.2, 8), (0.45, 8), (0.6, 3), (0.5, 1), (0.5, 32), (0.4, 40)])
def test_fill_holes_local_vs_labels(ops, density, area):
    """the bounded-flood-fill path (max_area <= 32) and the union-find path (> 32) against the reference
    statement (label everything, threshold the areas), incl. holes touching the border and the exact-area edge."""
    x = rnd(2, 1, 128, 96, seed=11)
    x = torch.where(torch.rand(x.shape, generator=gen(12)).cuda() < density, -x.abs(), x.abs())
    x[0, 0, :3, :3] = -1.0            # a 9-pixel component in the corner
    x[0, 0, 3, :4] = 1.0
    x[0, 0, :4, 3] = 1.0
    x[1, 0, 64, 10:18] = -1.0         # an exact 8-pixel bar ...
    x[1, 0, 63, 9:19] = 1.0
    x[1, 0, 65, 9:19] = 1.0
    x[1, 0, 64, 9] = 1.0
    x[1, 0, 64, 18] = 1.0
    assert torch.equal(ops.fill_holes(x, area), ref_ops.fill_holes(x, area))


# ------------------------------------------------------------------ norm / gemm
@pytest.mark.parametrize("M,C", [(1000, 96), (333, 768), (4096, 256), (5000, 4), (777, 16), (129, 64), (9, 256), (5001, 96),
                                 (4097, 192), (4099, 128)])
def test_layernorm(ops, M, C):
    x, a = rnd(M, C, seed=1, scale=2.0), rnd(M, C, seed=2)
    g, b = rnd(C, seed=3), rnd(C, seed=4)
    close(ops.layernorm(x, g, b, 1e-6), ref_ops.layernorm(x, g, b, 1e-6), 2e-5, "ln")
    close(ops.layernorm(x, g, b, 1e-5, add=a, act=1), ref_ops.layernorm(x, g, b, 1e-5, add=a, act=1), 2e-5, "ln+add+gelu")
    y = ops.layernorm(x, g, b, 1e-6, out_dtype=torch.bfloat16)
    close(y, ref_ops.layernorm(x, g, b, 1e-6), 5e-2, "ln bf16")


@pytest.mark.parametrize("M,N,K", [(65, 96, 96), (1000, 288, 96), (4096, 256, 64), (37, 4, 9), (513, 1024, 256),
                                   (9, 256, 2048), (262144 // 64, 16, 36), (2, 32, 256)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gemm_simt(ops, M, N, K, dt):
    a = rnd(M, K, seed=1).to(dt)
    w = (rnd(N, K, seed=2) / K ** 0.5).to(dt)
    bias, cs, res = rnd(N, seed=3), rnd(N, seed=4), rnd(M, N, seed=5)
    tol = 1e-4 if dt == torch.float32 else 2e-4
    close(ops.gemm(a, w, bias, impl=1), ref_ops.gemm(a, w, bias), tol, "gemm")
    close(ops.gemm(a, w, bias, act=1, residual=res, colscale=cs, impl=1),
          ref_ops.gemm(a, w, bias, act=1, residual=res, colscale=cs), tol, "gemm epilogue")
    close(ops.gemm(a, w, bias, act=2, out_dtype=torch.bfloat16, impl=1), ref_ops.gemm(a, w, bias, act=2), 3e-2, "gemm bf16 out")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (65, 96, 96), (1000, 288, 96), (4096, 256, 64),
                                   (513, 1024, 256), (4096, 2048, 256), (4096, 256, 2048), (65536, 288, 96),
                                   (16384, 576, 192), (4100, 1152, 384), (1024, 3072, 768), (300, 32, 256),
                                   (777, 64, 264), (129, 8, 16), (64, 40, 72)])
def test_gemm_tc(ops, M, N, K):
    """tcgen05 path (impl=2) vs the fp32 statement on the same bf16-rounded operands."""
    a = rnd(M, K, seed=1).to(torch.bfloat16)
    w = (rnd(N, K, seed=2) / K ** 0.5).to(torch.bfloat16)
    bias, cs, res = rnd(N, seed=3), rnd(N, seed=4), rnd(M, N, seed=5)
    close(ops.gemm(a, w, bias, impl=2), ref_ops.gemm(a, w, bias), 3e-4, "gemm_tc")
    close(ops.gemm(a, w, None, impl=2), ref_ops.gemm(a, w, None), 3e-4, "gemm_tc nobias")
    close(ops.gemm(a, w, bias, act=1, residual=res, colscale=cs, impl=2),
          ref_ops.gemm(a, w, bias, act=1, residual=res, colscale=cs), 3e-4, "gemm_tc epilogue")
    close(ops.gemm(a, w, bias, act=2, out_dtype=torch.bfloat16, impl=2), ref_ops.gemm(a, w, bias, act=2), 3e-2,
          "gemm_tc bf16 out")


def test_gemm_tc_strided_rows(ops):
    x = rnd(2, 4096, 768, seed=1).to(torch.bfloat16)
    w, b = (rnd(256, 256, seed=2) / 16).to(torch.bfloat16), rnd(256, seed=3)
    a = x.view(8192, 768)[:, 256:512]
    close(ops.gemm(a, w, b, impl=2), ref_ops.gemm(a.contiguous(), w, b), 3e-4, "strided A (column slice)")


@pytest.mark.parametrize("M,N,K", [(9, 256, 2048), (9, 2048, 256), (1, 4, 256), (3, 32, 256), (27, 256, 256), (64, 40, 72),
                                   (16, 1, 256), (10, 256, 128)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gemm_smallm(ops, M, N, K, dt):
    a = rnd(M, K, seed=1).to(dt)
    w = (rnd(N, K, seed=2) / K ** 0.5).to(dt)
    bias, cs, res = rnd(N, seed=3), rnd(N, seed=4), rnd(M, N, seed=5)
    tol = 1e-4 if dt == torch.float32 else 2e-4
    close(ops.gemm(a, w, bias, impl=3), ref_ops.gemm(a, w, bias), tol, "gemm smallm")
    close(ops.gemm(a, w, bias, act=1, residual=res, colscale=cs, impl=3),
          ref_ops.gemm(a, w, bias, act=1, residual=res, colscale=cs), tol, "gemm smallm epilogue")
    close(ops.gemm(a, w, bias, act=3, out_dtype=torch.bfloat16, impl=3), ref_ops.gemm(a, w, bias, act=3), 3e-2, "smallm bf16 out")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_gemm_grouped(ops, dt):
    """6 small-M GEMMs (different N / activation, row-strided A) in one launch == 6 separate statements."""
    x = rnd(3, 9, 256, seed=1).to(dt)
    a = [x.view(27, 256)[i::9] for i in range(6)]
    Ns, acts = [256, 256, 32, 4, 1, 40], [2, 0, 2, 3, 0, 1]
    ws = [(rnd(n, 256, seed=10 + i) / 16).to(dt) for i, n in enumerate(Ns)]
    bs = [rnd(n, seed=20 + i) for i, n in enumerate(Ns)]
    outs = ops.gemm_grouped(a, ws, bs, out_dtype=torch.float32, acts=acts)
    for o, ai, w, b, act in zip(outs, a, ws, bs, acts):
        close(o, ref_ops.gemm(ai.contiguous(), w, b, act=act), 1e-4 if dt == torch.float32 else 2e-4, "grouped")


def test_gemm_smallm_strided_rows(ops):
    x = rnd(4, 9, 256, seed=1).to(torch.bfloat16)
    w, b = (rnd(32, 256, seed=2) / 16).to(torch.bfloat16), rnd(32, seed=3)
    a = x.view(36, 256)[3::9]
    close(ops.gemm(a, w, b, impl=3), ref_ops.gemm(a.contiguous(), w, b), 2e-4, "strided A smallm")


@pytest.mark.parametrize("B,H,Lq,Lk,D", [(2, 8, 9, 4096, 16), (2, 8, 4096, 9, 16), (2, 8, 9, 9, 32), (1, 8, 16, 1024, 16),
                                         (3, 8, 1, 4096, 16), (1, 8, 1024, 32, 16), (2, 4, 7, 300, 32)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_attention_small(ops, B, H, Lq, Lk, D, dt):
    """decoder token-side attention kernels (impl=3)."""
    q, k, v = rnd(B, Lq, H * D, seed=1).to(dt), rnd(B, Lk, H * D, seed=2).to(dt), rnd(B, Lk, H * D, seed=3).to(dt)
    tol = 2e-5 if dt == torch.float32 else 1.5e-2
    close(ops.attention(q, k, v, H, impl=3), ref_ops.attention(q.float(), k.float(), v.float(), H), tol, "attention small")


def test_gemm_strided_rows(ops):
    x = rnd(4, 9, 256, seed=1)
    w, b = rnd(32, 256, seed=2) / 16, rnd(32, seed=3)
    a = x.view(36, 256)[3::9]
    close(ops.gemm(a, w, b, impl=1), ref_ops.gemm(a, w, b), 1e-4, "strided A")


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(2, 8, 9, 9, 32), (2, 8, 9, 1024, 16), (2, 8, 1024, 9, 16), (1, 1, 1024, 2056, 256),
                                         (1, 4, 1024, 1024, 96), (2, 1, 300, 77, 256), (1, 2, 5, 130, 64)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_attention_dense(ops, B, H, Lq, Lk, D, dt):
    q, k, v = rnd(B, Lq, H * D, seed=1).to(dt), rnd(B, Lk, H * D, seed=2).to(dt), rnd(B, Lk, H * D, seed=3).to(dt)
    tol = 2e-5 if dt == torch.float32 else 1.5e-2
    close(ops.attention(q, k, v, H, impl=1), ref_ops.attention(q.float(), k.float(), v.float(), H), tol, "attention")


@pytest.mark.parametrize("B,H,Lq,Lk,D,qs", [(1, 1, 128, 64, 256, 1.0), (1, 1, 256, 128, 64, 1.0), (1, 1, 128, 128, 128, 1.0),
                                            (1, 1, 128, 128, 96, 1.0), (1, 1, 4096, 4096 + 37, 256, 1.0),
                                            (1, 4, 4096, 4096, 96, 1.0), (2, 1, 300, 777, 256, 3.0),
                                            (1, 2, 1024, 1024, 64, 2.0), (2, 2, 256, 640, 128, 1.0),
                                            (1, 1, 4096, 28736, 256, 4.0), (3, 8, 200, 333, 96, 1.0),
                                            (1, 1, 128, 448, 256, 4.0), (1, 1, 256, 448, 256, 4.0), (1, 1, 300, 777, 256, 3.0),
                                            (1, 2, 640, 1200, 128, 4.0), (2, 4, 500, 900, 96, 5.0), (1, 1, 256, 4096, 64, 6.0)])
def test_attention_tc(ops, B, H, Lq, Lk, D, qs):
    """tcgen05 flash attention (impl=2, incl. split-KV + lazy rescale) vs the fp32 statement on the same
    bf16-rounded operands.  Tolerance 1.5e-2 abs: bf16 P and bf16 output rounding of O(1) values."""
    q = (rnd(B, Lq, H * D, seed=1) * qs).to(torch.bfloat16)
    k, v = rnd(B, Lk, H * D, seed=2).to(torch.bfloat16), rnd(B, Lk, H * D, seed=3).to(torch.bfloat16)
    o = ops.attention(q, k, v, H, impl=2)
    r = ref_ops.attention(q.float(), k.float(), v.float(), H)
    close_rel(o, r, 1e-2, 8e-3, "attention_tc")
    assert (o.float() - r).abs().mean().item() < 2e-3


@pytest.mark.parametrize("B,Lq,Lk,qs", [(1, 128, 64, 1.0), (1, 4096, 4096 + 37, 1.0), (2, 300, 777, 3.0), (1, 4096, 28736, 4.0)])
def test_attention_dv(ops, B, Lq, Lk, qs):
    """D=256 queries/keys over 64-d values (memory cross-attention with un-projected memory values)."""
    q = (rnd(B, Lq, 256, seed=1) * qs).to(torch.bfloat16)
    k, v = rnd(B, Lk, 256, seed=2), rnd(B, Lk, 64, seed=3).to(torch.bfloat16)
    if qs == 3.0:     # key norm ramps x9 along the bank: the row maximum outgrows the lazy-rescale threshold repeatedly
        k = k * (1 + 8 * torch.arange(Lk, device=k.device)[None, :, None] / Lk)
    k = k.to(torch.bfloat16)
    o = ops.attention_dv(q, k, v)
    r = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
    close_rel(o, r, 1e-2, 8e-3, "attention_dv")
    assert (o.float() - r).abs().mean().item() < 2e-3


@pytest.mark.parametrize("Lk,qs", [(209120, 1.0), (209120, 4.0), (1052736, 2.0)])
def test_attention_dv_benchmarked_key_counts(ops, Lk, qs):
    """memory cross-attention at the key counts of the benchmarked configurations: 209 120 = the longest bank of
    BASELINE configs[2] as bench.py runs it (48 conditioning memories, the recent memories, pointer tokens), 1 052 736 = configs[4] (256 conditioning memories).  Checked against an fp32 torch statement evaluated in query
    chunks on the same bf16-rounded operands; rows of probabilities over 10^5..10^6 keys make the outputs small, so the
    bound is relative to the output scale as well as absolute."""
    Lq = 4096
    q = (rnd(1, Lq, 256, seed=1) * qs).to(torch.bfloat16)
    k, v = rnd(1, Lk, 256, seed=2).to(torch.bfloat16), rnd(1, Lk, 64, seed=3).to(torch.bfloat16)
    # a few keys every query attends to strongly, so the softmax is not flat noise
    k[0, ::4099] = (q[0, :1].float() * 0.5).to(torch.bfloat16)
    o = ops.attention_dv(q, k, v).float()
    kf, vf = k[0].float(), v[0].float()
    r = torch.empty(Lq, 64, device="cuda")
    for c0 in range(0, Lq, 512):
        s = (q[0, c0:c0 + 512].float() @ kf.T) / 16.0
        r[c0:c0 + 512] = torch.softmax(s, dim=-1) @ vf
        del s
    err = (o[0] - r).abs()
    scale = r.abs().max().item()
    assert err.max().item() <= 1e-2 * max(scale, 1.0) and err.max().item() <= 8e-3 * scale + 2e-3, (err.max().item(), scale)
    assert err.mean().item() <= 2e-3 * max(r.abs().mean().item(), 1e-3) + 2e-4, (err.mean().item(), r.abs().mean().item())
    from conftest import record_parity
    record_parity(f"attention_dv_Lk{Lk}_qs{qs}", {"max_abs_err": err.max().item(), "out_abs_max": scale,
                                                  "mean_abs_err": err.mean().item(), "out_abs_mean": r.abs().mean().item()})


@pytest.mark.parametrize("B,Lq,cuts,qs", [(1, 4096, (64, 4197, 24576), 1.0), (1, 4096, (8192,), 4.0), (2, 256, (100, 300, 1000), 3.0)])
def test_attention_dv_partial_and_merge(ops, B, Lq, cuts, qs):
    """split-KV over several GPUs, on one GPU: partials over disjoint key ranges (plus an empty share) merged by
    ms2_attention_merge == attention over all keys."""
    Lk = cuts[-1]
    q = (rnd(B, Lq, 256, seed=1) * qs).to(torch.bfloat16)
    k = rnd(B, Lk, 256, seed=2).to(torch.bfloat16)
    v = rnd(B, Lk, 64, seed=3).to(torch.bfloat16)
    parts, lo = [], 0
    for hi in cuts:
        parts.append(ops.attention_dv_partial(q, k[:, lo:hi].contiguous(), v[:, lo:hi].contiguous()))
        lo = hi
    parts.append(ops.attention_dv_partial(q, None, None))
    parts = torch.stack(parts)
    o = ops.attention_merge(parts, B, Lq)
    ref = ref_ops.attention_dv(q, k, v).float()
    close_rel(o, ref, 1e-2, 8e-3, "partial+merge vs full attention")
    close_rel(o, ops.attention_dv(q, k, v), 1e-2, 8e-3, "partial+merge vs single-GPU kernel")
    # the partial of one share alone, merged, is plain attention over that share
    one = ops.attention_merge(ops.attention_dv_partial(q, k[:, : cuts[0]].contiguous(), v[:, : cuts[0]].contiguous())[None], B, Lq)
    close_rel(one, ref_ops.attention_dv(q, k[:, : cuts[0]], v[:, : cuts[0]]).float(), 1e-2, 8e-3, "single share")


def test_attention_tc_rescale_every_tile(ops, monkeypatch):
    """MS2_LAZY_TAU=0 forces the in-TMEM rescale of O whenever a row maximum grows (normally only beyond 2^8):
    exercises the softmax <-> MMA hand-shake of that rare path on every tile."""
    import subprocess, sys, os
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from medsam2_b200 import ops; import ref_ops\n"
        "g = torch.Generator().manual_seed(5)\n"
        "for (B,H,Lq,Lk,D,qs) in [(1,1,300,777,256,3.0),(1,1,4096,28736,256,4.0),(2,4,500,900,96,4.0)]:\n"
        "    q=(torch.randn(B,Lq,H*D,generator=g)*qs).cuda().bfloat16(); k=torch.randn(B,Lk,H*D,generator=g).cuda().bfloat16(); v=torch.randn(B,Lk,H*D,generator=g).cuda().bfloat16()\n"
        "    o=ops.attention(q,k,v,H,impl=2); r=ref_ops.attention(q.float(),k.float(),v.float(),H)\n"
        "    e=((o.float()-r).abs()-8e-3*r.abs()).max().item(); assert e<=1e-2,(B,H,Lq,Lk,D,e)\n"
        # memory cross-attention (64-d values): with tau = 0 the speculative exponentials of a tile are redone whenever
        # the row maximum grew; the second case ramps the key norm so the maximum keeps growing along the bank
        "import torch.nn.functional as F\n"
        "for (B,Lq,Lk,qs,ramp) in [(1,4096,28736,4.0,0.0),(2,300,5000,2.0,6.0),(1,256,200,1.0,0.0)]:\n"
        "    q=(torch.randn(B,Lq,256,generator=g)*qs).cuda().bfloat16(); k=torch.randn(B,Lk,256,generator=g).cuda()\n"
        "    k=(k*(1+ramp*torch.arange(Lk,device='cuda')[None,:,None]/Lk)).bfloat16(); v=torch.randn(B,Lk,64,generator=g).cuda().bfloat16()\n"
        "    o=ops.attention_dv(q,k,v); r=F.scaled_dot_product_attention(q.float()[:,None],k.float()[:,None],v.float()[:,None])[:,0]\n"
        "    e=((o.float()-r).abs()-8e-3*r.abs()).max().item(); assert e<=1e-2,('dv',B,Lq,Lk,e)\n"
        "print('ok')\n") % (os.path.join(os.path.dirname(__file__), "..", "medical-sam2_b200"), os.path.dirname(__file__))
    env = dict(os.environ, MS2_LAZY_TAU="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_attention_tc_strided_qkv(ops):
    B, L, C = 2, 512, 256
    qkv = rnd(B, L, 3 * C, seed=7).to(torch.bfloat16)
    o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], 1, impl=2)
    r = ref_ops.attention(qkv[:, :, :C].float().contiguous(), qkv[:, :, C:2 * C].float().contiguous(),
                          qkv[:, :, 2 * C:].float().contiguous(), 1)
    close(o, r, 1.5e-2, "strided qkv tc")
    qkv = rnd(1, 1024, 3 * 384, seed=8).to(torch.bfloat16)      # Hiera global block layout: 4 heads of 96
    o = ops.attention(qkv[:, :, :384], qkv[:, :, 384:768], qkv[:, :, 768:], 4, impl=2)
    r = ref_ops.attention(qkv[:, :, :384].float().contiguous(), qkv[:, :, 384:768].float().contiguous(),
                          qkv[:, :, 768:].float().contiguous(), 4)
    close(o, r, 1.5e-2, "strided qkv tc d96")


def test_attention_strided_qkv(ops):
    B, L, C = 2, 256, 256
    qkv = rnd(B, L, 3 * C, seed=7)
    o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], 1, impl=1)
    r = ref_ops.attention(qkv[:, :, :C].contiguous(), qkv[:, :, C:2 * C].contiguous(), qkv[:, :, 2 * C:].contiguous(), 1)
    close(o, r, 2e-5, "strided qkv")


@pytest.mark.parametrize("H,W,heads,ws,qpool", [(64, 64, 1, 8, 0), (64, 64, 2, 8, 1), (32, 32, 2, 4, 0), (32, 32, 4, 4, 1),
                                                (64, 64, 4, 14, 0), (64, 64, 8, 14, 1), (32, 32, 8, 7, 0), (20, 36, 2, 14, 0)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_window_attention(ops, H, W, heads, ws, qpool, dt):
    B, D = 2, 96
    qkv = rnd(B, H, W, 3 * heads * D, seed=1).to(dt)
    bias = rnd(3 * heads * D, seed=2)
    tol = 2e-5 if dt == torch.float32 else 1.5e-2
    close(ops.window_attention(qkv, bias, B, H, W, heads, D, ws, qpool, impl=1),
          ref_ops.window_attention(qkv.float(), bias, B, H, W, heads, D, ws, qpool), tol, "window attention")


@pytest.mark.parametrize("H,W,heads,ws,qpool", [(64, 64, 1, 8, 0), (64, 64, 2, 8, 1), (32, 32, 2, 4, 0), (32, 32, 4, 4, 1),
                                                (64, 64, 4, 14, 0), (64, 64, 8, 14, 1), (32, 32, 8, 7, 0), (20, 36, 2, 14, 0),
                                                (256, 256, 1, 8, 0), (24, 40, 2, 8, 1), (12, 20, 4, 4, 0), (30, 18, 8, 7, 0)])
def test_window_attention_tc(ops, H, W, heads, ws, qpool):
    """tcgen05 windowed attention (impl=2): packed windows, pad-token-as-bias keys, q-pool, crop."""
    B, D = 2, 96
    qkv = rnd(B, H, W, 3 * heads * D, seed=1).to(torch.bfloat16)
    bias = rnd(3 * heads * D, seed=2)
    o = ops.window_attention(qkv, bias, B, H, W, heads, D, ws, qpool, impl=2)
    # the reference statement sees the bias rounded to bf16 exactly as the kernel's pad tokens do
    r = ref_ops.window_attention(qkv.float(), bias.to(torch.bfloat16).float(), B, H, W, heads, D, ws, qpool)
    close_rel(o, r, 1e-2, 8e-3, "window attention tc")
    assert (o.float() - r).abs().mean().item() < 2e-3


# ------------------------------------------------------------------ conv-shaped / elementwise
def test_patch_embed(ops):
    img = rnd(2, 3, 256, 320, seed=1)
    w, b, pos = rnd(96, 3, 7, 7, seed=2) / 12, rnd(96, seed=3), rnd(64, 80, 96, seed=4)
    close(ops.patch_embed(img, w, b, pos), ref_ops.patch_embed(img, w, b, pos), 1e-4, "patch_embed")


@pytest.mark.parametrize("k,s,p,Cin,pre", [(3, 2, 1, 1, 1), (3, 2, 1, 1, 2), (3, 2, 1, 4, 0), (3, 2, 1, 16, 0), (2, 2, 0, 1, 0),
                                           (2, 2, 0, 4, 0), (4, 4, 0, 1, 0)])
def test_im2col(ops, k, s, p, Cin, pre):
    x = rnd(2, 64, 48, Cin, seed=1)
    close(ops.im2col(x, k, s, p, torch.float32, pre, 20.0 if pre else 1.0, -10.0 if pre else 0.0),
          ref_ops.im2col(x, k, s, p, torch.float32, pre, 20.0 if pre else 1.0, -10.0 if pre else 0.0), 1e-5, "im2col")


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 256), (1, 13, 21, 40), (3, 8, 8, 32), (1, 5, 70, 96)])
def test_dwconv7x7(ops, B, H, W, C):
    x, w, b = rnd(B, H, W, C, seed=1), rnd(C, 49, seed=2) / 7, rnd(C, seed=3)
    close(ops.dwconv7x7(x, w, b), ref_ops.dwconv7x7(x, w, b), 1e-4, "dwconv")
    close(ops.dwconv7x7(x, w, None), ref_ops.dwconv7x7(x, w, None), 1e-4, "dwconv nobias")


@pytest.mark.parametrize("cin,cout,pre", [(1, 4, 1), (1, 4, 2), (4, 16, 0), (16, 64, 0)])
def test_conv3x3s2_ln_gelu(ops, cin, cout, pre):
    x = rnd(2, 36, 50, cin, seed=1)
    w, b = rnd(cout, cin, 3, 3, seed=2) / 3, rnd(cout, seed=3)
    g, be = rnd(cout, seed=4), rnd(cout, seed=5)
    ps, pb = (20.0, -10.0) if pre else (1.0, 0.0)
    close(ops.conv3x3s2_ln_gelu(x, w, b, g, be, 1e-6, pre=pre, pre_scale=ps, pre_bias=pb),
          ref_ops.conv3x3s2_ln_gelu(x, w, b, g, be, 1e-6, pre=pre, pre_scale=ps, pre_bias=pb), 2e-4, "conv3x3s2+ln+gelu")
    close(ops.conv3x3s2_ln_gelu(x, w, b, g, be, 1e-6, out_dtype=torch.bfloat16),
          ref_ops.conv3x3s2_ln_gelu(x, w, b, g, be, 1e-6), 3e-2, "conv3x3s2+ln+gelu bf16")


def test_im2col_vec8(ops):
    x = rnd(2, 32, 24, 64, seed=1)
    close(ops.im2col(x, 3, 2, 1, torch.bfloat16), ref_ops.im2col(x, 3, 2, 1, torch.bfloat16), 1e-6, "im2col vec8")


def test_patch_im2col_gemm_matches_conv(ops):
    """bf16 patch-embed path: im2col rows x re-laid-out weight == conv 7x7/s4/p3 on the same bf16-rounded operands."""
    img = rnd(2, 3, 96, 128, seed=1)
    w, b = rnd(96, 3, 7, 7, seed=2) / 12, rnd(96, seed=3)
    cols = ops.patch_im2col(img)
    w152 = F.pad(w.permute(0, 2, 3, 1).reshape(96, -1), (0, 5)).to(torch.bfloat16).contiguous()
    y = ops.gemm(cols, w152, b)
    ref = F.conv2d(img.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b, stride=4, padding=3).permute(0, 2, 3, 1)
    close(y.view(2, 24, 32, 96), ref, 3e-4, "patch im2col+gemm")


def test_elementwise_family(ops):
    x, z = rnd(3, 100, 64, seed=1), rnd(3, 100, 64, seed=2)
    close(ops.axpby(x, 1.0, z, 0.1), ref_ops.axpby(x, 1.0, z, 0.1), 1e-6, "axpby")
    close(ops.axpby(x, 20.0, None, 0.0, -10.0), ref_ops.axpby(x, 20.0, None, 0.0, -10.0), 1e-5, "axpby const")
    close(ops.axpby(x, 1.0, z[0], 1.0), ref_ops.axpby(x, 1.0, z[0], 1.0), 1e-6, "axpby bcast")
    close(ops.axpby(x, 1.0, z, 1.0, out_dtype=torch.bfloat16), ref_ops.axpby(x, 1.0, z, 1.0), 3e-2, "axpby bf16")
    xo, zo = rnd(3, 7, 5, seed=6), rnd(3, 7, 5, seed=7)          # odd sizes: the scalar kernels
    close(ops.axpby(xo, 2.0, zo, 0.5, 1.0), ref_ops.axpby(xo, 2.0, zo, 0.5, 1.0), 1e-6, "axpby odd")
    close(ops.axpby(xo, 1.0, zo[0], 1.0, out_dtype=torch.bfloat16), ref_ops.axpby(xo, 1.0, zo[0], 1.0), 3e-2, "axpby odd bcast")
    close(ops.cast(ops.cast(xo, torch.bfloat16), torch.float32), xo.bfloat16().float(), 0.0, "cast odd")
    close(ops.add_rowvec(xo, rnd(5, seed=8)), ref_ops.add_rowvec(xo, rnd(5, seed=8)), 1e-6, "add_rowvec odd")
    v = rnd(64, seed=3)
    close(ops.add_rowvec(x, v), ref_ops.add_rowvec(x, v), 1e-6, "add_rowvec")
    close(ops.cast(ops.cast(x, torch.bfloat16), torch.float32), x.bfloat16().float(), 0.0, "cast")
    for act in (1, 2, 3):
        close(ops.activation(x, act), ref_ops.activation(x, act), 1e-6, f"act {act}")
    gate = torch.tensor([1.0, -1.0, 0.0]).cuda()
    close(ops.gate_rows(x, gate, -1024.0), ref_ops.gate_rows(x, gate, -1024.0), 0.0, "gate_rows")
    idx = torch.tensor([2, 0, 1], dtype=torch.int32).cuda()
    xs = rnd(3, 4, 16, 16, seed=5)
    close(ops.select_plane(xs, idx), ref_ops.select_plane(xs, idx), 0.0, "select_plane")
    cnt = ops.mask_stability_counts(xs, 0.05)
    assert torch.equal(cnt, ref_ops.mask_stability_counts(xs, 0.05))


def test_layout_kernels(ops):
    x = rnd(2, 32, 48, 96, seed=1)
    close(ops.maxpool2x2(x), ref_ops.maxpool2x2(x), 0.0, "maxpool")
    xs_ = rnd(1, 6, 10, 6, seed=21)                                   # C % 4 != 0: the scalar kernel
    close(ops.maxpool2x2(xs_), ref_ops.maxpool2x2(xs_), 0.0, "maxpool scalar")
    close(ops.nhwc_to_nchw(x), ref_ops.nhwc_to_nchw(x), 0.0, "nhwc->nchw")
    y = rnd(2, 96, 32, 48, seed=2)
    close(ops.nchw_to_nhwc(y), ref_ops.nchw_to_nhwc(y), 0.0, "nchw->nhwc")
    fine, coarse = rnd(2, 32, 48, 96, seed=3), rnd(2, 16, 24, 96, seed=4)
    close(ops.upsample2x_add_(fine.clone(), coarse), ref_ops.upsample2x_add_(fine.clone(), coarse), 0.0, "upsample2x_add")
    g, bias, skip = rnd(2, 16, 16, 4 * 64, seed=5), rnd(64, seed=6), rnd(2, 32, 32, 64, seed=7)
    close(ops.pixel_shuffle_add(g, bias, skip, 2, 16, 16, 64, act=1), ref_ops.pixel_shuffle_add(g, bias, skip, 2, 16, 16, 64, act=1),
          1e-6, "pixel_shuffle_add")
    g2, skip2 = rnd(1, 9, 7, 4 * 6, seed=10), rnd(1, 18, 14, 6, seed=11)         # C % 4 != 0: the scalar kernel, no bias
    close(ops.pixel_shuffle_add(g2, None, skip2, 1, 9, 7, 6), ref_ops.pixel_shuffle_add(g2, None, skip2, 1, 9, 7, 6), 1e-6,
          "pixel_shuffle_add scalar")
    up, hyper = rnd(2, 4096, 32, seed=8), rnd(2, 4, 32, seed=9)
    close(ops.hyper_mask(up, hyper), ref_ops.hyper_mask(up, hyper), 1e-4, "hyper_mask")
    up, hyper = rnd(2, 1000, 32, seed=12), rnd(2, 4, 32, seed=13)                  # ragged pixel count
    close(ops.hyper_mask(up, hyper), ref_ops.hyper_mask(up, hyper), 1e-4, "hyper_mask ragged")
    up, hyper = rnd(1, 333, 16, seed=14), rnd(1, 3, 16, seed=15)                   # generic (warp-per-pixel) kernel
    close(ops.hyper_mask(up, hyper), ref_ops.hyper_mask(up, hyper), 1e-4, "hyper_mask generic")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_rope(ops, dt):
    from medsam2_b200.modeling.position_encoding import rope_table
    cos, sin = rope_table(256, 16, 16, 10000.0, "cuda")
    B, L, n_ptr = 2, 256 * 3 + 8, 8
    x = rnd(B, L, 256, seed=1).to(dt)
    r = ref_ops.rope_(x.clone(), B, L, L - n_ptr, 256, cos, sin)
    y = ops.rope_(x.clone(), B, L, L - n_ptr, 256, cos, sin)
    close(y, r, 1e-6 if dt == torch.float32 else 2e-2, "rope")
    # strided (q slice of a fused qkv buffer)
    qkv = rnd(B, 256, 768, seed=2)
    r2 = qkv.clone()
    r2[:, :, 256:512] = ref_ops.rope_(qkv[:, :, 256:512].contiguous(), B, 256, 256, 256, cos, sin)
    y2 = qkv.clone()
    ops.rope_(y2[:, :, 256:], B, 256, 256, 256, cos, sin, batch_stride=256 * 768, row_stride=768)
    close(y2, r2, 1e-6, "rope strided")


@pytest.mark.parametrize("size,out,aa", [((256, 256), (1024, 1024), 0), ((128, 128), (512, 512), 0), ((256, 256), (333, 517), 0),
                                         ((1024, 1024), (256, 256), 1), ((700, 500), (256, 256), 1), ((1024, 1024), (256, 256), 0)])
def test_resize(ops, size, out, aa):
    x = rnd(3, *size, seed=1)
    close(ops.resize_bilinear(x, out, aa), ref_ops.resize_bilinear(x, out, aa), 2e-5, "resize")


def test_point_embed(ops):
    g = rnd(2, 64, seed=1)
    table = rnd(5, 128, seed=2)
    coords = (rnd(3, 4, 2, seed=3).abs() * 300).contiguous()
    labels = torch.tensor([[1, 0, 2, 3], [-1, 1, 1, 0], [2, 3, -1, -1]], dtype=torch.int32).cuda()
    for pad in (True, False):
        close(ops.point_embed(coords, labels, g, table, pad, (1024, 768)), ref_ops.point_embed(coords, labels, g, table, pad, (1024, 768)),
              2e-4, f"point_embed pad={pad}")
    close(ops.point_embed(coords, labels.long(), g, table, True, (512, 512)), ref_ops.point_embed(coords, labels, g, table, True, (512, 512)),
          2e-4, "point_embed int64 labels")


def test_fourier_and_normalize(ops):
    c, gm = torch.rand(5, 3, 2, generator=gen(1)).cuda(), rnd(2, 128, seed=2)
    close(ops.fourier_pe(c, gm), ref_ops.fourier_pe(c, gm), 2e-5, "fourier_pe")
    x = (torch.rand(2, 3, 64, 64, generator=gen(3)) * 255).cuda()
    close(ops.normalize_image(x), ref_ops.normalize_image(x), 1e-5, "normalize f32")
    u = (torch.rand(2, 64, 64, 3, generator=gen(4)) * 255).to(torch.uint8).cuda()
    close(ops.normalize_image(u), ref_ops.normalize_image(u), 1e-5, "normalize u8")


# ------------------------------------------------------------------ post-processing
@pytest.mark.parametrize("shape", [(1, 1, 64, 64), (2, 1, 48, 40), (5, 1, 48, 40), (13, 1, 1024, 1024), (3, 1, 33, 7)])
def test_non_overlap_bit_exact(ops, shape):
    """ms2_non_overlap vs the torch statement of sam2_base.py:812-830 (argmax keeps the FIRST object on ties)."""
    x = rnd(*shape, seed=3, scale=4.0)
    x[:, :, :4] = x[:1, :, :4]                                  # ties across all objects
    got = ops.non_overlap(x)
    assert torch.equal(got, ref_ops.non_overlap(x)), shape


def test_non_overlap_reference_golden(ops):
    z = np.load(f"{G}/postprocess_cases.npz")
    for n in (1, 2, 5):
        x = torch.from_numpy(z[f"no/in_{n}"]).cuda()
        from medsam2_b200.modeling.sam2_base import SAM2Base
        got = SAM2Base._apply_non_overlapping_constraints(None, x)
        assert np.array_equal(got.cpu().numpy(), z[f"no/out_{n}"]), n


@pytest.mark.parametrize("N,h,H,T", [(3, 64, 256, 5), (2, 256, 1024, 5), (1, 128, 512, 8), (4, 50, 120, 1), (96, 256, 1024, 5)])
def test_score_lowres_equals_resize_then_score(ops, N, h, H, T):
    """ms2_score_lowres (bilinear up-sampling fused into the eval_seg counts + BCE pass) == ms2_resize_bilinear followed
    by ms2_seg_counts (bit-exact integers) and ms2_bce_logits_sum (fp64 sums of fp32 terms, 1e-6 relative)."""
    low = rnd(N, h, h, seed=4, scale=0.6)
    low[0, : h // 4] = 0.1                                      # hole-fill value exactly on a threshold
    gt = (rnd(N, H, H, seed=5) > 0.3).float()
    thr = [0.1, 0.3, 0.5, 0.7, 0.9, -0.2, 0.0, 0.05][:T]
    counts, sums = ops.score_lowres(low, gt, thr, pos_weight=2.0)
    up = ops.resize_bilinear(low, (H, H))
    assert torch.equal(counts, ops.seg_counts(up, gt, thr))
    want = ops.bce_logits_sum(up, gt, 2.0)
    assert ((sums - want).abs() <= 1e-6 * want.abs()).all(), (sums, want)
    c2, s2 = ops.score_lowres(low, gt, thr)
    assert s2 is None and torch.equal(c2, counts)
    # and the up-sampling itself against torch (fp32 rounding only)
    close(up, F.interpolate(low[:, None], size=(H, H), mode="bilinear", align_corners=False)[:, 0], 2e-6, "bilinear")


# ------------------------------------------------------------------ per-slice glue kernels (csrc/glue.cu)
@pytest.mark.parametrize("B,dt", [(1, torch.bfloat16), (2, torch.bfloat16), (1, torch.float32)])
def test_bank_rows(ops, B, dt):
    """recent memories + pointer tokens, concatenated / position-coded / cast in one launch == the torch statement;
    destinations may be row ranges of a larger bank; > 80 sources are chunked."""
    W = 64
    srcs = [rnd(B, 4096, W, seed=i) for i in range(3)] + [rnd(B, 4, W, seed=10 + i) for i in range(5)]
    table = rnd(4096, W, seed=20)
    per_b = rnd(B, 4, W, seed=21)
    poss = [table, None, table, per_b, None, None, per_b.expand(B, 4, W)[:1].expand(B, 4, W), None]
    total = sum(t.shape[1] for t in srcs)
    bank = torch.zeros(B, total + 100, W, dtype=dt, device="cuda")
    k, m = ops.bank_rows(srcs, poss, dt, m_out=bank[:, 37:37 + total])
    rk, rm = ref_ops.bank_rows(srcs, poss, dt, m_out=torch.empty(B, total, W, dtype=dt, device="cuda"))
    assert torch.equal(k, rk) and torch.equal(bank[:, 37:37 + total], rm)
    assert not bank[:, :37].any() and not bank[:, 37 + total:].any()
    many = [rnd(B, 4, W, seed=100 + i) for i in range(173)]
    k2, _ = ops.bank_rows(many, [None] * 173, dt)
    assert torch.equal(k2, torch.cat(many, dim=1).to(dt))


def test_sam_head_glue(ops):
    B, M, C = 5, 3, 256
    ious = rnd(B, 4, seed=1)
    ious[1, 1] = ious[1, 2] = 9.0                                    # tie: the first maximum wins
    hs = rnd(B, 9, C, seed=2)
    idx, rows = ops.argmax_select_rows(ious[:, 1:], hs[:, 2:5])      # strided views, as the decoder hands them over
    ridx, rrows = ref_ops.argmax_select_rows(ious[:, 1:], hs[:, 2:5])
    assert torch.equal(idx, ridx) and torch.equal(rows, rrows)
    idx2, none = ops.argmax_select_rows(ious.contiguous())
    assert none is None and torch.equal(idx2, torch.argmax(ious, -1).int())
    ptr, logits, no_obj = rnd(B, C, seed=3), torch.tensor([3.0, -2.0, 0.0, 1e-3, -7.0]).cuda(), rnd(C, seed=4)
    for soft in (False, True):
        for fixed in (False, True):
            close(ops.obj_ptr_mix(ptr, logits, no_obj, soft, fixed), ref_ops.obj_ptr_mix(ptr, logits, no_obj, soft, fixed),
                  2e-6, f"obj_ptr_mix soft={soft} fixed={fixed}")
    counts = torch.tensor([[98, 100], [99, 100], [0, 0], [5, 10], [100, 100]], dtype=torch.int32).cuda()
    i3, iou3 = ops.stability_select(counts, ious.contiguous(), 0.98)
    r3, riou3 = ref_ops.stability_select(counts, ious.contiguous(), 0.98)
    assert torch.equal(i3, r3) and torch.equal(iou3, riou3)
    # point embedding with constant rows in front (the decoder's output tokens)
    gauss, table, prefix = rnd(2, 128, seed=5), rnd(5, 256, seed=6), rnd(6, 256, seed=7)
    coords = (torch.rand(B, 2, 2, generator=gen(8)) * 1024).cuda()
    labels = torch.tensor([[2, 3]] * B, dtype=torch.int32).cuda()
    close(ops.point_embed(coords, labels, gauss, table, True, (1024, 1024), prefix),
          ref_ops.point_embed(coords, labels, gauss, table, True, (1024, 1024), prefix), 2e-4, "point_embed + prefix")


def test_multi_copy(ops):
    """several device-to-device copies in one launch (graph runner inputs / result clones): dense tensors of mixed
    dtypes, sizes from 4 B to 8 MB, unaligned storage offsets, a channels-last view, more than 16 items, an empty one;
    a strided (non-dense) pair takes the copy_ fall-back."""
    g = torch.Generator(device="cuda").manual_seed(3)
    srcs = [torch.randn(n, device="cuda", generator=g) for n in (1, 7, 1024, 2 * 1024 * 1024, 333)]
    srcs += [torch.randint(0, 255, (n,), device="cuda", dtype=torch.uint8, generator=g) for n in (5, 4099)]
    srcs.append(torch.randn(1000, device="cuda", generator=g)[3:])                    # 12-byte storage offset
    srcs.append(torch.randn(2, 8, 16, 16, device="cuda", generator=g).permute(0, 3, 1, 2))   # dense, not contiguous
    srcs += [torch.randn(64 + i, device="cuda", generator=g).to(torch.bfloat16) for i in range(20)]
    srcs.append(torch.empty(0, device="cuda"))
    srcs.append(torch.randn(16, 16, device="cuda", generator=g)[:, ::2])              # not dense -> copy_
    dsts = [torch.empty_like(s) if i != len(srcs) - 1 else torch.empty(16, 8, device="cuda") for i, s in enumerate(srcs)]
    ops.multi_copy(list(zip(dsts, srcs)))
    for d, s in zip(dsts, srcs):
        assert torch.equal(d, s)


@pytest.mark.parametrize("M,K,out_dtype", [(4096, 256, torch.bfloat16), (4096, 64, torch.bfloat16), (4096, 2048, torch.bfloat16),
                                           (128, 256, torch.float32), (8192, 2048, torch.float32)])
def test_gemm_res_ln(ops, M, K, out_dtype):
    """x = residual + A W^T + b and t = LayerNorm(x) in one kernel (memory_attention.py:58-99,166) vs the fp32 statement
    on the same bf16-rounded operands: x to fp32 accumulation-order accuracy, t to the output rounding."""
    a = rnd(M, K, seed=1).to(torch.bfloat16)
    w = (rnd(256, K, seed=2) / K ** 0.5).to(torch.bfloat16)
    b, g, be = rnd(256, seed=3), 1 + 0.1 * rnd(256, seed=4), 0.1 * rnd(256, seed=5)
    res = rnd(M, 256, seed=6) * 3 + 5                         # a mean well away from zero
    x_ref = res + a.float() @ w.float().T + b
    t_ref = F.layer_norm(x_ref, (256,), g, be, 1e-5)
    x, t = ops.gemm_res_ln(a, w, b, res.clone(), g, be, 1e-5, out_dtype=out_dtype, x_out=torch.empty_like(res))
    assert t.dtype == out_dtype
    close(x, x_ref, 2e-5 * max(1.0, K ** 0.5 / 8), "gemm_res_ln x")
    close(t, t_ref, 1e-4 if out_dtype == torch.float32 else 2e-2, "gemm_res_ln t")
    # in place over the residual, and the same LayerNorm as the stand-alone kernel on the same x
    r2 = res.clone()
    x2, t2 = ops.gemm_res_ln(a, w, b, r2, g, be, 1e-5, out_dtype=out_dtype)
    assert x2.data_ptr() == r2.data_ptr() and torch.equal(x2, x)
    close(t2, ops.layernorm(x, g, be, 1e-5, out_dtype=out_dtype), 1e-5 if out_dtype == torch.float32 else 3.2e-2, "vs ms2_layernorm")   # one bf16 ulp of |t| < 8


@pytest.mark.parametrize("B,L,N,rope_cols,K", [(1, 4096, 256, 256, 256), (1, 4096, 768, 512, 256), (2, 1024, 768, 512, 256),
                                              (1, 128, 256, 0, 64)])
def test_gemm_rope(ops, B, L, N, rope_cols, K):
    """projection + rotary encoding in one kernel == ms2_gemm (bf16 out) followed by ms2_rope on the rotated column
    tiles, bit for bit (the kernel rounds the projection to bf16 before rotating, like the two-kernel path)."""
    a = rnd(B * L, K, seed=1).to(torch.bfloat16)
    w = (rnd(N, K, seed=2) / K ** 0.5).to(torch.bfloat16)
    b = rnd(N, seed=3)
    T = min(L, 4096)
    ang = rnd(T, 128, seed=4) * 3
    cos_t, sin_t = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
    got = ops.gemm_rope(a, w, b, L, rope_cols, cos_t, sin_t)
    ref = ops.gemm(a, w, b, out_dtype=torch.bfloat16)
    for c0 in range(0, rope_cols, 256):
        ops.rope_(ref[:, c0:], B, L, L, 256, cos_t, sin_t, batch_stride=L * N, row_stride=N)
    assert torch.equal(got, ref), (got.float() - ref.float()).abs().max().item()


@pytest.mark.parametrize("B,H,W,dt", [(1, 1024, 1024, torch.float32), (2, 512, 520, torch.bfloat16), (1, 28, 300, torch.float32)])
def test_patch_im2col_tiled(ops, B, H, W, dt):
    """the shared-memory tiled patch gather (64 output pixels per CTA, partial last tile, bf16 frames) == the statement"""
    img = rnd(B, 3, H, W, seed=5).to(dt)
    got = ops.patch_im2col(img)
    ref = ref_ops.patch_im2col(img.float())
    assert got.shape == ref.shape and torch.equal(got.float(), ref.float())


def test_to_device_async_ring():
    """prompt tensors staged through the pinned ring: values survive slot reuse (more uploads than slots, no host
    synchronisation in between), oversize tensors take the ordinary path, device tensors pass through"""
    from medsam2_b200.utils.misc import _PinnedRing, to_device_async
    outs, refs = [], []
    for i in range(3 * _PinnedRing.SLOTS + 5):
        t = torch.full((1, 1 + i % 7, 2), float(i)) + torch.arange(2)
        lab = torch.full((1, 1 + i % 7), i, dtype=torch.int32)
        outs += [to_device_async(t, "cuda"), to_device_async(lab, "cuda")]
        refs += [t, lab]
    big = torch.arange(1000, dtype=torch.float32)
    assert torch.equal(to_device_async(big, "cuda").cpu(), big)
    d = torch.ones(3, device="cuda")
    assert to_device_async(d, "cuda") is d
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert o.is_cuda and o.dtype == r.dtype and torch.equal(o.cpu(), r)
