"""Tensor-level wrappers over the C-ABI kernels (include/medsam2_b200.h).

PyTorch is used here for device memory (torch.empty), streams and nothing else: every function
below validates its tensors, hands raw device pointers + the current CUDA stream to the native
library and raises if the call fails.  There is no eager/PyTorch fallback.
"""
import math

import torch

from . import native

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
_DT = {torch.float32: F32, torch.bfloat16: BF16}
_SMS = 148                  # B200; only sizes split-KV scratch (the kernels query the device themselves)
# debugging aid: MS2_DISABLE=tc_gemm,small_gemm,tc_attn,small_attn,tc_win forces the fp32-accumulate SIMT
# kernel for that family (still a CUDA kernel of this library - there is no non-native path)
import os as _os
_DISABLED = set(filter(None, _os.environ.get("MS2_DISABLE", "").split(",")))


_raw_stream = torch._C._cuda_getCurrentRawStream     # (device_index) -> cudaStream_t as int, ~0.3 us
_cur_dev = torch._C._cuda_getDevice                 # (torch.cuda.current_device without the lazy-init checks: 0.1 us)


def _st():
    return _raw_stream(_cur_dev())


class _EventProfiler:
    """Optional per-launch CUDA-event timing of selected ops on the launching stream (bench.py's
    roofline numbers).  Off by default: zero overhead on the product path."""

    def __init__(self):
        self.names = None
        self.records = []

    def enable(self, names):
        self.names = set(names)
        self.records = []

    def disable(self):
        self.names = None

    def begin(self, name):
        if self.names is None or name not in self.names:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name, e0, flops):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((name, flops, e0, e1))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, flops, e0, e1 in self.records:
            d = out.setdefault(name, {"ms": 0.0, "flops": 0.0, "n": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["n"] += 1
        return out


PROFILE = _EventProfiler()


def _chk(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise native.NativeError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.device.index != _cur_dev():
        raise native.NativeError(f"{name}: tensor lives on cuda:{t.device.index} but the current device is cuda:{_cur_dev()} "
                                 "(kernels launch on the current device's stream: call torch.cuda.set_device first)")
    if not t.is_contiguous():
        raise native.NativeError(f"{name}: tensor must be contiguous, got strides {t.stride()}")
    if dtype is not None and t.dtype != dtype:
        raise native.NativeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _opt(t, name, dtype=torch.float32):
    return None if t is None else _chk(t, name, dtype)


def dt_of(t):
    return _DT[t.dtype]


# ------------------------------------------------------------------ connected components
def cc_label(mask_u8):
    """[N,1,H,W] uint8 -> (labels, counts) int32 (reference `_C.get_connected_componnets`)."""
    if mask_u8.dim() != 4 or mask_u8.shape[1] != 1:
        raise native.NativeError("inputs must be [N, 1, H, W] shape")
    N, _, H, W = mask_u8.shape
    p = _chk(mask_u8, "inputs", torch.uint8)
    labels = torch.empty((N, 1, H, W), dtype=torch.int32, device=mask_u8.device)
    counts = torch.empty_like(labels)
    ws = None
    if (H // 2) * (W // 2) > 28672:
        ws = torch.empty((N, H, W), dtype=torch.int32, device=mask_u8.device)
    native.call("ms2_cc_label", p, labels.data_ptr(), counts.data_ptr(), None if ws is None else ws.data_ptr(),
                N, H, W, _st())
    return labels, counts


def fill_holes(scores, max_area, thresh=0.0, fill_value=0.1):
    N, _, H, W = scores.shape
    out = torch.empty_like(scores)
    native.call("ms2_fill_holes", _chk(scores, "scores", torch.float32), out.data_ptr(), N, H, W, float(thresh),
                int(max_area), float(fill_value), _st())
    return out


# ------------------------------------------------------------------ norm / gemm
def layernorm(x, gamma, beta, eps, out_dtype=torch.float32, add=None, act=ACT_NONE):
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    native.call("ms2_layernorm", _chk(x, "x", torch.float32), _opt(add, "add"), _chk(gamma, "gamma", torch.float32),
                _chk(beta, "beta", torch.float32), y.data_ptr(), _DT[out_dtype], M, C, float(eps), act, _st())
    return y


def gemm(a, w, bias=None, out_dtype=torch.float32, act=ACT_NONE, residual=None, colscale=None, impl=0, out=None):
    """out[..., N] = residual + colscale * act(a[..., K] @ w[N, K]^T + bias)."""
    K = a.shape[-1]
    N = w.shape[0]
    if w.shape[1] != K:
        raise native.NativeError(f"gemm: K mismatch {a.shape} x {w.shape}")
    M = a.numel() // K
    lda = K
    if a.dim() == 2 and a.stride(1) == 1 and a.stride(0) != K and a.is_cuda:
        lda = a.stride(0)              # row-strided A (e.g. one token column of [B, Nt, C])
        a_ptr = a.data_ptr()
    else:
        a_ptr = _chk(a, "a")
    if out is None:
        out = torch.empty(a.shape[:-1] + (N,), dtype=out_dtype, device=a.device)
    if impl == 0 and _DISABLED:
        if M >= 64 and "tc_gemm" in _DISABLED:
            impl = 1
        if M < 64 and "small_gemm" in _DISABLED:
            impl = 1
    ev = PROFILE.begin("gemm")
    native.call("ms2_gemm", a_ptr, _DT[a.dtype], lda, _chk(w, "w"), _DT[w.dtype], _opt(bias, "bias"),
                _opt(colscale, "colscale"), _opt(residual, "residual"), N, _chk(out, "out"), _DT[out.dtype], N,
                M, N, K, act, impl, _st())
    PROFILE.end("gemm", ev, 2.0 * M * N * K)
    return out


def gemm_grouped(a_list, w_list, bias_list, out_dtype=torch.float32, acts=None, outs=None):
    """len(a_list) <= 8 independent small-M GEMMs in one launch: out_g = act_g(a_g[M,K] @ w_g[N_g,K]^T + bias_g).
    a_g may be row-strided 2-D views (token rows of [B,Nt,C]); all share M, K and dtype.  `outs`: optional
    destinations (row-strided [M,N_g] views of `out_dtype`, None entries are allocated)."""
    import ctypes
    G = len(a_list)
    M, K = a_list[0].shape
    dst = list(outs) if outs is not None else [None] * G
    outs, A, W, Bs, O, lda, ldo, Ns = [], [], [], [], [], [], [], []
    for a, w, b, o in zip(a_list, w_list, bias_list, dst):
        if a.shape != (M, K) or w.shape[1] != K or a.stride(1) != 1 or not a.is_cuda or a.dtype != a_list[0].dtype:
            raise native.NativeError("gemm_grouped: operands must share M, K, dtype and be CUDA with unit inner stride")
        if o is None:
            o = torch.empty((M, w.shape[0]), dtype=out_dtype, device=a.device)
        elif tuple(o.shape) != (M, w.shape[0]) or o.dtype != out_dtype or o.stride(1) != 1 or not o.is_cuda:
            raise native.NativeError("gemm_grouped: bad destination view")
        outs.append(o)
        A.append(a.data_ptr()); W.append(_chk(w, "w")); Bs.append(0 if b is None else _chk(b, "bias", torch.float32))
        O.append(o.data_ptr()); lda.append(a.stride(0)); ldo.append(o.stride(0)); Ns.append(w.shape[0])
    acts = list(acts) if acts is not None else [ACT_NONE] * G
    vp, lp, ip = ctypes.c_void_p * G, ctypes.c_long * G, ctypes.c_int * G
    native.call("ms2_gemm_smallm_grouped", G, vp(*A), lp(*lda), vp(*W), vp(*Bs), vp(*O), lp(*ldo), ip(*Ns), ip(*acts),
                _DT[a_list[0].dtype], _DT[out_dtype], M, K, _st())
    return outs


# ------------------------------------------------------------------ attention
def attention(q, k, v, heads, scale=None, impl=0):
    """q [B,Lq,heads*D], k/v [B,Lk,heads*D] (token-major, possibly column slices of a wider
    matrix via as_strided views with last-dim stride 1) -> o [B,Lq,heads*D] contiguous."""
    B, Lq, HD = q.shape
    Lk = k.shape[1]
    D = HD // heads
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.stride(-1) != 1:
            raise native.NativeError(f"attention: {n} must be a CUDA tensor with unit inner stride")
    o = torch.empty((B, Lq, HD), dtype=q.dtype, device=q.device)
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    if impl == 0 and _DISABLED:
        if D in (16, 32) and "small_attn" in _DISABLED:
            impl = 1
        if D >= 64 and "tc_attn" in _DISABLED:
            impl = 1
    ws, ws_bytes = None, 0
    if impl != 1 and q.dtype == torch.bfloat16 and D in (64, 96, 128, 256) and Lq >= 64 and Lk >= 512:
        qtiles = B * heads * ((Lq + 127) // 128)
        if qtiles < 2 * _SMS:                                   # too few query tiles to fill the SMs: split-KV scratch
            per_split = B * heads * Lq * (D + 2) * 4
            nsplit = max(2, min(32, (4 * _SMS) // qtiles, (256 << 20) // per_split))
            ws_bytes = nsplit * per_split
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    if impl in (0, 3) and D in (16, 32) and Lq <= 16 and Lk >= 512:
        ws_bytes = B * heads * 64 * Lq * (D + 2) * 4             # decoder token->image attention: key-split partials
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    ev = PROFILE.begin("attention")
    native.call("ms2_attention_ws", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _DT[q.dtype],
                q.stride(0), D, q.stride(1), k.stride(0), D, k.stride(1), v.stride(0), D, v.stride(1),
                o.stride(0), D, o.stride(1), B, heads, Lq, Lk, D, float(scale), impl,
                None if ws is None else ws.data_ptr(), ws_bytes, _st())
    PROFILE.end("attention", ev, 4.0 * B * heads * Lq * Lk * D)
    return o


def attention_dv(q, k, v, scale=None):
    """single-head attention with a narrower value: q [B,Lq,D], k [B,Lk,D], v [B,Lk,DV] -> o [B,Lq,DV]
    (tcgen05 only; D=256, DV=64: memory cross-attention over un-projected memory values)."""
    B, Lq, D = q.shape
    Lk, DV = k.shape[1], v.shape[2]
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.stride(-1) != 1 or t.dtype != torch.bfloat16:
            raise native.NativeError(f"attention_dv: {n} must be a bf16 CUDA tensor with unit inner stride")
    o = torch.empty((B, Lq, DV), dtype=q.dtype, device=q.device)
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    ws, ws_bytes = None, 0
    qtiles = B * ((Lq + 127) // 128)
    if qtiles < 2 * _SMS and Lk >= 512:
        per_split = B * Lq * (DV + 2) * 4
        nsplit = max(2, min(32, (4 * _SMS) // qtiles, (256 << 20) // per_split))
        ws_bytes = nsplit * per_split
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    ev = PROFILE.begin("mem_cross_attention")
    native.call("ms2_attention_dv", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _DT[q.dtype],
                q.stride(0), D, q.stride(1), k.stride(0), D, k.stride(1), v.stride(0), DV, v.stride(1),
                o.stride(0), DV, o.stride(1), B, 1, Lq, Lk, D, DV, float(scale),
                None if ws is None else ws.data_ptr(), ws_bytes, _st())
    PROFILE.end("mem_cross_attention", ev, 2.0 * B * Lq * Lk * (D + DV))
    return o


def attention_dv_partial(q, k, v, out=None, scale=None):
    """this rank's share of a split-KV memory cross-attention: q [B,Lq,256], k [B,Lk,256], v [B,Lk,64] (bf16) ->
    packed fp32 partial [B*Lq*(64+2)]: B*Lq*64 un-normalised outputs followed by B*Lq (row max in log2 units, row sum)
    pairs.  k = None or Lk == 0 gives the empty partial (0, -inf, 0)."""
    B, Lq, D = q.shape
    DV = 64
    rows = B * Lq
    if out is None:
        out = torch.empty(rows * (DV + 2), dtype=torch.float32, device=q.device)
    part_o, part_ml = out[: rows * DV], out[rows * DV:]
    Lk = 0 if k is None else k.shape[1]
    if Lk == 0:
        part_o.zero_()
        ml = part_ml.view(rows, 2)
        ml[:, 0] = float("-inf")
        ml[:, 1] = 0.0
        return out
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.stride(-1) != 1 or t.dtype != torch.bfloat16:
            raise native.NativeError(f"attention_dv_partial: {n} must be a bf16 CUDA tensor with unit inner stride")
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    ws, ws_bytes = None, 0
    qtiles = B * ((Lq + 127) // 128)
    if qtiles < 2 * _SMS and Lk >= 512:
        per_split = rows * (DV + 2) * 4
        nsplit = max(2, min(32, (4 * _SMS) // qtiles, (256 << 20) // per_split))
        ws_bytes = nsplit * per_split
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    ev = PROFILE.begin("mem_cross_attention")
    native.call("ms2_attention_dv_partial", q.data_ptr(), k.data_ptr(), v.data_ptr(), part_o.data_ptr(), part_ml.data_ptr(),
                _DT[q.dtype], q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), B, Lq, Lk, D,
                v.shape[2], float(scale), None if ws is None else ws.data_ptr(), ws_bytes, _st())
    PROFILE.end("mem_cross_attention", ev, 2.0 * B * Lq * Lk * (D + DV))
    return out


def attention_merge(parts, B, Lq, DV=64):
    """parts fp32 [nparts, B*Lq*(DV+2)] (packed partials of attention_dv_partial, one per rank) -> o bf16 [B,Lq,DV]."""
    nparts = parts.shape[0]
    rows = B * Lq
    _chk(parts, "parts", torch.float32)
    assert parts.shape[1] == rows * (DV + 2)
    o = torch.empty((B, Lq, DV), dtype=torch.bfloat16, device=parts.device)
    native.call("ms2_attention_merge", parts.data_ptr(), parts.data_ptr() + rows * DV * 4, parts.stride(0), o.data_ptr(),
                _DT[o.dtype], o.stride(0), o.stride(1), B, Lq, DV, nparts, _st())
    return o


def attention_dv_partial_push(q, k, v, dst_ptrs, flag_ptrs, step, counter, scale=None):
    """This rank's share of a split-KV memory cross-attention, pushed into EVERY rank's gather buffer over peer
    memory by the kernel that folds the local splits (no collective call): q [B,Lq,256], k [B,Lk,256], v [B,Lk,64]
    bf16 (k = None or Lk == 0: the empty partial).  dst_ptrs[r] / flag_ptrs[r]: device address of this rank's slot /
    flag word in rank r's symmetric memory; `counter`: zeroed int32 [1] on this device."""
    import ctypes
    B, Lq, D = q.shape
    DV = 64
    rows = B * Lq
    world = len(dst_ptrs)
    Lk = 0 if k is None else k.shape[1]
    if Lk:
        for t, n in ((q, "q"), (k, "k"), (v, "v")):
            if not t.is_cuda or t.stride(-1) != 1 or t.dtype != torch.bfloat16:
                raise native.NativeError(f"attention_dv_partial_push: {n} must be a bf16 CUDA tensor with unit inner stride")
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    per_split = rows * (DV + 2) * 4
    qtiles = B * ((Lq + 127) // 128)
    nsplit = max(2, min(32, (4 * _SMS) // qtiles, (256 << 20) // per_split)) if Lk >= 512 else 1
    ws = torch.empty(nsplit * per_split, dtype=torch.uint8, device=q.device)
    ev = PROFILE.begin("mem_cross_attention")
    native.call("ms2_attention_dv_partial_push", q.data_ptr(), None if not Lk else k.data_ptr(), None if not Lk else v.data_ptr(),
                BF16, q.stride(0), q.stride(1), 0 if not Lk else k.stride(0), 0 if not Lk else k.stride(1),
                0 if not Lk else v.stride(0), 0 if not Lk else v.stride(1), B, Lq, Lk, D, DV, float(scale), ws.data_ptr(),
                ws.numel(), (ctypes.c_void_p * world)(*dst_ptrs), (ctypes.c_void_p * world)(*flag_ptrs), world, int(step),
                _chk(counter, "counter", torch.int32), _st())
    PROFILE.end("mem_cross_attention", ev, 2.0 * B * Lq * Lk * (D + DV))


def attention_merge_wait(parts, flags, step, B, Lq, DV=64):
    """parts fp32 [world, B*Lq*(DV+2)] (this rank's gather buffer, filled by the peers' attention_dv_partial_push),
    flags int32 [world]: waits on the device until every flag shows `step`, then merges -> bf16 [B,Lq,DV]."""
    world = parts.shape[0]
    o = torch.empty((B, Lq, DV), dtype=torch.bfloat16, device=parts.device)
    native.call("ms2_attention_merge_wait", _chk(parts, "parts", torch.float32), parts.stride(0), _chk(flags, "flags", torch.int32),
                int(step), world, o.data_ptr(), BF16, o.stride(0), o.stride(1), B, Lq, DV, _st())
    return o


def window_attention(qkv, qkv_bias, B, H, W, heads, D, ws, qpool, impl=0):
    """qkv [B,H,W,3*heads*D] -> [B,Ho,Wo,heads*D]."""
    Ho, Wo = (H // 2, W // 2) if qpool else (H, W)
    out = torch.empty((B, Ho, Wo, heads * D), dtype=qkv.dtype, device=qkv.device)
    if impl == 0 and "tc_win" in _DISABLED:
        impl = 1
    ev = PROFILE.begin("window_attention")
    native.call("ms2_window_attention_impl", _chk(qkv, "qkv"), _chk(qkv_bias, "qkv_bias", torch.float32),
                out.data_ptr(), _DT[qkv.dtype], B, H, W, heads, D, ws, int(bool(qpool)), 1.0 / math.sqrt(D), impl, _st())
    nwin = B * ((H + ws - 1) // ws) * ((W + ws - 1) // ws)
    lq = (ws // 2) ** 2 if qpool else ws * ws
    PROFILE.end("window_attention", ev, 4.0 * nwin * heads * lq * ws * ws * D)
    return out


# ------------------------------------------------------------------ element-wise / layout
def maxpool2x2(x):
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), dtype=torch.float32, device=x.device)
    native.call("ms2_maxpool2x2", _chk(x, "x", torch.float32), y.data_ptr(), B, H, W, C, _st())
    return y


def patch_embed(img, w, bias, pos):
    B, _, Hin, Win = img.shape
    Cout = w.shape[0]
    Ho, Wo = (Hin + 6 - 7) // 4 + 1, (Win + 6 - 7) // 4 + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=img.device)
    native.call("ms2_patch_embed", _chk(img, "img", torch.float32), _chk(w, "w", torch.float32),
                _chk(bias, "bias", torch.float32), _opt(pos, "pos"), out.data_ptr(), B, Hin, Win, Cout, _st())
    return out


def patch_im2col(img, ldk=152):
    """fp32 or bf16 NCHW image [B,3,H,W] -> bf16 [B, Ho*Wo, ldk] rows of 7x7/s4/p3 taps ((ky,kx,c) order, zero padded)."""
    B, _, Hin, Win = img.shape
    Ho, Wo = (Hin + 6 - 7) // 4 + 1, (Win + 6 - 7) // 4 + 1
    cols = torch.empty((B, Ho * Wo, ldk), dtype=torch.bfloat16, device=img.device)
    native.call("ms2_patch_im2col", _chk(img, "img"), _DT[img.dtype], cols.data_ptr(), B, Hin, Win, ldk, _st())
    return cols


def axpby(x, a=1.0, z=None, b=1.0, c=0.0, out_dtype=torch.float32, out=None):
    """y = a*x + b*z + c ; z broadcasts as z.flatten()[i mod z.numel()] (trailing-dims broadcast)."""
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device) if out is None else out
    native.call("ms2_axpby", _chk(x, "x", torch.float32), float(a), _opt(z, "z"), float(b), float(c), y.data_ptr(),
                _DT[y.dtype], x.numel(), 0 if z is None else z.numel(), _st())
    return y


def gate_rows(x, gate, fill):
    """x fp32 [B, ...]; gate fp32 [B] (or [B,1]) -> where(gate>0, x, fill)."""
    B = x.shape[0]
    y = torch.empty_like(x)
    native.call("ms2_gate_rows", _chk(x, "x", torch.float32), _chk(gate, "gate", torch.float32), float(fill),
                y.data_ptr(), B, x.numel() // max(B, 1), _st())
    return y


def select_plane(x, idx):
    """x fp32 [B, M, ...], idx int32 [B] -> [B, 1, ...] = x[b, idx[b]]."""
    B, M = x.shape[:2]
    P = x.numel() // max(B * M, 1)
    y = torch.empty((B, 1) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    native.call("ms2_select_plane", _chk(x, "x", torch.float32), _chk(idx, "idx", torch.int32), y.data_ptr(),
                B, M, P, _st())
    return y


def add_rowvec(x, v, s=1.0):
    C = x.shape[-1]
    y = torch.empty_like(x)
    native.call("ms2_add_rowvec", _chk(x, "x", torch.float32), _chk(v, "v", torch.float32), float(s), y.data_ptr(),
                x.numel() // C, C, _st())
    return y


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    native.call("ms2_cast", _chk(x, "x"), _DT[x.dtype], y.data_ptr(), _DT[dtype], x.numel(), _st())
    return y


def cast_into(x, out):
    """dtype cast of contiguous x into the (contiguous) pre-allocated `out`."""
    if x.numel() != out.numel():
        raise native.NativeError("cast_into: size mismatch")
    native.call("ms2_cast", _chk(x, "x"), _DT[x.dtype], _chk(out, "out"), _DT[out.dtype], x.numel(), _st())
    return out


def activation(x, act):
    y = torch.empty_like(x)
    native.call("ms2_activation", _chk(x, "x", torch.float32), y.data_ptr(), x.numel(), act, _st())
    return y


def upsample2x_add_(fine, coarse):
    B, H, W, C = fine.shape
    native.call("ms2_upsample2x_add", _chk(fine, "fine", torch.float32), _chk(coarse, "coarse", torch.float32),
                B, H, W, C, _st())
    return fine


def nhwc_to_nchw(x):
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    native.call("ms2_nhwc_to_nchw", _chk(x, "x", torch.float32), y.data_ptr(), B, H, W, C, _st())
    return y


def nchw_to_nhwc(x):
    B, C, H, W = x.shape
    y = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
    native.call("ms2_nchw_to_nhwc", _chk(x, "x", torch.float32), y.data_ptr(), B, C, H, W, _st())
    return y


def rope_(x, B, rows, n_rope_rows, D, cos_t, sin_t, batch_stride=None, row_stride=None):
    """in place on x viewed as [B, rows, D] with the given element strides."""
    if not x.is_cuda:
        raise native.NativeError("rope: expected a CUDA tensor")
    native.call("ms2_rope", x.data_ptr(), _DT[x.dtype], batch_stride if batch_stride is not None else rows * D,
                row_stride if row_stride is not None else D, B, rows, n_rope_rows, D,
                _chk(cos_t, "cos", torch.float32), _chk(sin_t, "sin", torch.float32), cos_t.shape[0], _st())
    return x


def im2col(x, k, stride, pad, out_dtype, pre=0, pre_scale=1.0, pre_bias=0.0):
    B, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    cols = torch.empty((B, Ho, Wo, k * k * Cin), dtype=out_dtype, device=x.device)
    native.call("ms2_im2col", _chk(x, "x", torch.float32), cols.data_ptr(), _DT[out_dtype], B, H, W, Cin, k, stride,
                pad, pre, float(pre_scale), float(pre_bias), _st())
    return cols


def conv3x3s2_ln_gelu(x, w, bias, gamma, beta, eps, out_dtype=torch.float32, pre=0, pre_scale=1.0, pre_bias=0.0):
    """fused conv3x3/s2/p1 + LayerNorm2d + GELU on NHWC fp32 x [B,H,W,Cin]; w fp32 [Cout,Cin,3,3]."""
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=out_dtype, device=x.device)
    native.call("ms2_conv3x3s2_ln_gelu", _chk(x, "x", torch.float32), _chk(w, "w", torch.float32),
                _chk(bias, "bias", torch.float32), _chk(gamma, "gamma", torch.float32), _chk(beta, "beta", torch.float32),
                y.data_ptr(), _DT[out_dtype], B, H, W, Cin, Cout, float(eps), pre, float(pre_scale), float(pre_bias), _st())
    return y


def dwconv7x7(x, w, bias):
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    native.call("ms2_dwconv7x7", _chk(x, "x", torch.float32), _chk(w, "w", torch.float32), _opt(bias, "bias"),
                y.data_ptr(), B, H, W, C, _st())
    return y


def pixel_shuffle_add(g, bias, skip, B, H, W, C, act=ACT_NONE):
    out = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=g.device)
    native.call("ms2_pixel_shuffle_add", _chk(g, "g", torch.float32), _opt(bias, "bias"), _opt(skip, "skip"),
                out.data_ptr(), B, H, W, C, act, _st())
    return out


def hyper_mask(up, hyper):
    B, P, C = up.shape
    Mk = hyper.shape[1]
    masks = torch.empty((B, Mk, P), dtype=torch.float32, device=up.device)
    native.call("ms2_hyper_mask", _chk(up, "up", torch.float32), _chk(hyper, "hyper", torch.float32),
                masks.data_ptr(), B, P, C, Mk, _st())
    return masks


def resize_bilinear(x, size, antialias=False):
    """x fp32 [..., H, W] -> [..., Ho, Wo], align_corners=False."""
    H, W = x.shape[-2:]
    Ho, Wo = size
    N = x.numel() // (H * W)
    y = torch.empty(x.shape[:-2] + (Ho, Wo), dtype=torch.float32, device=x.device)
    native.call("ms2_resize_bilinear", _chk(x, "x", torch.float32), y.data_ptr(), N, H, W, Ho, Wo,
                int(bool(antialias)), _st())
    return y


def fourier_pe(coords01, gauss):
    n = coords01.numel() // 2
    Fh = gauss.shape[1]
    out = torch.empty(coords01.shape[:-1] + (2 * Fh,), dtype=torch.float32, device=coords01.device)
    native.call("ms2_fourier_pe", _chk(coords01, "coords", torch.float32), _chk(gauss, "gauss", torch.float32),
                out.data_ptr(), n, Fh, _st())
    return out


def point_embed(coords, labels, gauss, table, pad, image_size, prefix=None):
    """coords fp32 [B,N,2] (input pixels), labels int [B,N] -> sparse prompt embeddings fp32 [B,P+N+pad,2F]; `prefix`
    fp32 [P,2F]: constant rows written in front of the prompt rows of every batch entry."""
    B, N, _ = coords.shape
    Fh = gauss.shape[1]
    P = 0 if prefix is None else prefix.shape[0]
    out = torch.empty((B, P + N + (1 if pad else 0), 2 * Fh), dtype=torch.float32, device=coords.device)
    lab = labels if labels.dtype == torch.int32 else labels.to(torch.int32)
    cf = coords if (coords.dtype == torch.float32 and coords.is_contiguous()) else coords.float().contiguous()
    native.call("ms2_point_embed", _chk(cf, "coords", torch.float32),
                _chk(lab if lab.is_contiguous() else lab.contiguous(), "labels", torch.int32), _chk(gauss, "gauss", torch.float32),
                _chk(table, "table", torch.float32), out.data_ptr(), B, N, 1 if pad else 0, Fh, int(image_size[1]),
                int(image_size[0]), None if prefix is None else _chk(prefix, "prefix", torch.float32), P, _st())
    return out


def bank_rows(srcs, poss, out_dtype, k_out=None, m_out=None):
    """srcs: list of fp32 [B,rows_i,W] contiguous tensors; poss: per source None, an fp32 [rows_i,W] table or an fp32
    [B,rows_i,W] tensor.  Writes the sources back to back (in order) as k_out[b] = src + pos and m_out[b] = src, both
    `out_dtype` [B,total,W] (row-strided views of a larger buffer are fine: unit inner stride, rows contiguous);
    allocates k_out when None, m_out stays optional.  -> (k_out, m_out)."""
    import ctypes
    n = len(srcs)
    B, _, W = srcs[0].shape
    total = sum(t.shape[1] for t in srcs)
    if k_out is None:
        k_out = torch.empty((B, total, W), dtype=out_dtype, device=srcs[0].device)
    if n > 80:                                   # the kernel takes its source pointers by value: 80 per launch
        r0 = 0
        for c0 in range(0, n, 80):
            rows_c = sum(t.shape[1] for t in srcs[c0:c0 + 80])
            bank_rows(srcs[c0:c0 + 80], poss[c0:c0 + 80], out_dtype, k_out[:, r0:r0 + rows_c],
                      None if m_out is None else m_out[:, r0:r0 + rows_c])
            r0 += rows_c
        return k_out, m_out
    for d in (k_out, m_out):
        if d is not None and (tuple(d.shape) != (B, total, W) or d.dtype != out_dtype or d.stride(2) != 1 or d.stride(1) != W):
            raise native.NativeError("bank_rows: destination must be [B,total,W] with contiguous rows")
    src_p, pos_p, pos_bs, rows = [], [], [], []
    for t, q in zip(srcs, poss):
        if t.shape[0] != B or t.shape[2] != W:
            raise native.NativeError("bank_rows: sources must share B and W")
        src_p.append(_chk(t, "src", torch.float32))
        rows.append(t.shape[1])
        if q is None:
            pos_p.append(None); pos_bs.append(0)
        else:
            if q.dim() == 3 and q.shape[0] == B and q.stride(0) != 0:
                pos_bs.append(q.stride(0))
            else:
                q = q[0] if q.dim() == 3 else q
                pos_bs.append(0)
            if q.stride(-1) != 1 or q.stride(-2) != W or tuple(q.shape[-2:]) != (t.shape[1], W) or q.dtype != torch.float32:
                raise native.NativeError("bank_rows: position tensor must be fp32 [rows,W] / [B,rows,W] with contiguous rows")
            pos_p.append(q.data_ptr())
    native.call("ms2_bank_rows", (ctypes.c_void_p * n)(*src_p), (ctypes.c_void_p * n)(*pos_p), (ctypes.c_long * n)(*pos_bs),
                (ctypes.c_int * n)(*rows), n, W, B, k_out.data_ptr(), k_out.stride(0),
                None if m_out is None else m_out.data_ptr(), 0 if m_out is None else m_out.stride(0), _DT[out_dtype], _st())
    return k_out, m_out


def argmax_select_rows(scores, rows=None):
    """scores fp32 [B,M] -> (idx int32 [B] = first arg-max, rows[b, idx[b]] fp32 [B,C] or None); rows [B,M,C] may be a
    strided view with unit inner stride."""
    B, M = scores.shape
    idx = torch.empty((B,), dtype=torch.int32, device=scores.device)
    out = None
    rp, bs, rs, C = None, 0, 0, 0
    if rows is not None:
        if rows.dtype != torch.float32 or rows.stride(-1) != 1 or rows.shape[0] != B or rows.shape[1] < M:
            raise native.NativeError("argmax_select_rows: rows must be fp32 [B,>=M,C] with unit inner stride")
        C = rows.shape[2]
        out = torch.empty((B, C), dtype=torch.float32, device=scores.device)
        rp, bs, rs = rows.data_ptr(), rows.stride(0), rows.stride(1)
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.stride(1) != 1:
        raise native.NativeError("argmax_select_rows: scores must be fp32 CUDA [B,M] with unit inner stride")
    native.call("ms2_argmax_select_rows", scores.data_ptr(), scores.stride(0), B, M, rp, bs, rs, C, idx.data_ptr(),
                None if out is None else out.data_ptr(), _st())
    return idx, out


def obj_ptr_mix(ptr, logits, no_obj, soft, fixed):
    """(fixed ? lam*ptr : ptr) + (1-lam)*no_obj with lam = sigmoid(logit) (soft) or (logit > 0)."""
    B, C = ptr.shape
    out = torch.empty_like(ptr)
    native.call("ms2_obj_ptr_mix", _chk(ptr, "ptr", torch.float32), _chk(logits, "logits", torch.float32),
                _chk(no_obj, "no_obj", torch.float32), out.data_ptr(), B, C, int(bool(soft)), int(bool(fixed)), _st())
    return out


def stability_select(counts, ious, thresh):
    """counts int32 [B,2] (mask_stability_counts), ious fp32 [B,M] -> (idx int32 [B], iou fp32 [B,1])."""
    B, M = ious.shape
    idx = torch.empty((B,), dtype=torch.int32, device=ious.device)
    iou = torch.empty((B, 1), dtype=torch.float32, device=ious.device)
    native.call("ms2_stability_select", _chk(counts, "counts", torch.int32), _chk(ious, "ious", torch.float32), B, M,
                float(thresh), idx.data_ptr(), iou.data_ptr(), _st())
    return idx, iou


def gemm_res_ln(a, w, bias, residual, gamma, beta, eps, out_dtype=torch.bfloat16, x_out=None):
    """x = residual + a @ w.T + bias (fp32, written to x_out or in place over residual) and t = LayerNorm(x):
    a bf16 [M,K], w bf16 [256,K], residual fp32 [M,256] -> (x fp32 [M,256], t out_dtype [M,256]); M % 128 == 0."""
    M, K = a.shape
    x = residual if x_out is None else x_out
    t = torch.empty((M, 256), dtype=out_dtype, device=a.device)
    native.call("ms2_gemm_res_ln", _chk(a, "a", torch.bfloat16), K, _chk(w, "w", torch.bfloat16), _opt(bias, "bias"),
                _chk(residual, "residual", torch.float32), 256, _chk(x, "x_out", torch.float32), 256,
                _chk(gamma, "gamma", torch.float32), _chk(beta, "beta", torch.float32), float(eps), t.data_ptr(), _DT[t.dtype],
                256, M, K, _st())
    return x, t


def gemm_rope(a, w, bias, L, rope_cols, cos_t, sin_t):
    """bf16 [M,N] = a @ w.T + bias with the 256-wide column tiles below rope_cols rotated (row r: position r mod L)."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    native.call("ms2_gemm_rope", _chk(a, "a", torch.bfloat16), K, _chk(w, "w", torch.bfloat16), _opt(bias, "bias"),
                out.data_ptr(), N, M, N, K, int(L), int(rope_cols), _chk(cos_t, "cos", torch.float32),
                _chk(sin_t, "sin", torch.float32), cos_t.shape[0], _st())
    return out


def _dense(t):
    """the tensor covers its storage span exactly once (contiguous up to a permutation of the dimensions)"""
    if t.is_contiguous():
        return True
    expect = 1
    for st, sz in sorted((st, sz) for sz, st in zip(t.shape, t.stride()) if sz != 1):
        if st != expect:
            return False
        expect *= sz
    return True


def multi_copy(pairs):
    """[(dst, src), ...]: dst.copy_(src) for every pair, in ONE launch per 16 pairs when both tensors of a pair are dense
    with the same dtype and strides (then the copy is a byte copy of the storage span); other pairs fall back to copy_."""
    import ctypes
    fast = []
    for d, s in pairs:
        if (d.dtype == s.dtype and d.shape == s.shape and d.stride() == s.stride() and d.is_cuda and s.is_cuda
                and d.device == s.device and _dense(d)):
            if d.numel():
                fast.append((d, s))
        else:
            d.copy_(s)
    if not fast:
        return
    n = len(fast)
    if n == 1:
        fast[0][0].copy_(fast[0][1])
        return
    srcs = (ctypes.c_void_p * n)(*[s.data_ptr() for _, s in fast])
    dsts = (ctypes.c_void_p * n)(*[d.data_ptr() for d, _ in fast])
    sizes = (ctypes.c_long * n)(*[d.numel() * d.element_size() for d, _ in fast])
    if fast[0][0].device.index != _cur_dev():
        raise native.NativeError("multi_copy: tensors live on another device than the current one")
    native.call("ms2_multi_copy", ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(dsts, ctypes.c_void_p),
                ctypes.cast(sizes, ctypes.c_void_p), n, _st())


def normalize_image(x, out=None, nhwc=None, out_dtype=torch.float32):
    """(x/255 - mean)/std -> NCHW `out_dtype` (fp32, or bf16 = the cast autocast applies before the patch-embed conv).
    x: fp32 [B,3,H,W] in 0..255 (video tensor), uint8 [B,3,H,W] (uint8 video tensor: `imgs_tensor / 255.0` of
    utils/misc.py:233 works for any dtype) or, with nhwc=True, uint8 [B,H,W,3] (decoded images).  nhwc=None infers the
    uint8 layout from the shape and refuses ambiguous or malformed ones instead of mis-reading them."""
    if x.dim() != 4:
        raise native.NativeError(f"normalize_image: expected a 4-D tensor, got {tuple(x.shape)}")
    if x.dtype == torch.uint8:
        if nhwc is None:
            c_first, c_last = x.shape[1] == 3, x.shape[-1] == 3
            if c_first == c_last:
                raise native.NativeError(f"normalize_image: cannot tell the layout of a uint8 tensor of shape "
                                         f"{tuple(x.shape)}: pass nhwc=True ([B,H,W,3]) or nhwc=False ([B,3,H,W])")
            nhwc = c_last
        if nhwc:
            if x.shape[-1] != 3:
                raise native.NativeError(f"normalize_image: uint8 NHWC input must be [B,H,W,3], got {tuple(x.shape)}")
            B, H, W, _ = x.shape
            layout = 1
        else:
            if x.shape[1] != 3:
                raise native.NativeError(f"normalize_image: uint8 NCHW input must be [B,3,H,W], got {tuple(x.shape)}")
            B, _, H, W = x.shape
            layout = 2
        _chk(x, "x", torch.uint8)
    else:
        if x.shape[1] != 3:
            raise native.NativeError(f"normalize_image: expected [B,3,H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        layout = 0
        _chk(x, "x", torch.float32)
    if out is None:
        out = torch.empty((B, 3, H, W), dtype=out_dtype, device=x.device)
    elif tuple(out.shape) != (B, 3, H, W):
        raise native.NativeError("normalize_image: bad `out` shape")
    native.call("ms2_normalize_image", x.data_ptr(), layout, _chk(out, "out"), _DT[out.dtype], B, H, W, _st())
    return out


def mask_stability_counts(x, delta):
    """x fp32 [N, ...] -> int32 [N,2] = (#>delta, #>-delta)."""
    N = x.shape[0]
    P = x.numel() // max(N, 1)
    counts = torch.empty((N, 2), dtype=torch.int32, device=x.device)
    native.call("ms2_mask_stability_counts", _chk(x, "x", torch.float32), counts.data_ptr(), N, P, float(delta), _st())
    return counts


def seg_counts(pred, gt, thresholds):
    """pred, gt fp32 [N, ...] (same shape) -> int32 [N,T,3] = (#(pred>th & gt>th), #(pred>th), #(gt>th)) per threshold."""
    import ctypes
    T = len(thresholds)
    if not 1 <= T <= 8:
        raise ValueError("seg_counts: 1..8 thresholds per call")
    if pred.shape != gt.shape:
        raise ValueError(f"seg_counts: pred {tuple(pred.shape)} vs gt {tuple(gt.shape)}")
    N = pred.shape[0]
    P = pred.numel() // max(N, 1)
    counts = torch.empty((N, T, 3), dtype=torch.int32, device=pred.device)
    thr = (ctypes.c_float * T)(*[float(t) for t in thresholds])
    native.call("ms2_seg_counts", _chk(pred, "pred", torch.float32), _chk(gt, "gt", torch.float32),
                ctypes.cast(thr, ctypes.c_void_p), T, counts.data_ptr(), N, P, _st())
    return counts


def bce_logits_sum(pred, gt, pos_weight):
    """pred (logits), gt fp32 [N, ...] -> fp64 [N]: per-plane sum of the BCE-with-logits element losses."""
    if pred.shape != gt.shape:
        raise ValueError(f"bce_logits_sum: pred {tuple(pred.shape)} vs gt {tuple(gt.shape)}")
    N = pred.shape[0]
    P = pred.numel() // max(N, 1)
    sums = torch.empty((N,), dtype=torch.float64, device=pred.device)
    native.call("ms2_bce_logits_sum", _chk(pred, "pred", torch.float32), _chk(gt, "gt", torch.float32), float(pos_weight),
                sums.data_ptr(), N, P, _st())
    return sums


def score_lowres(low, gt, thresholds, pos_weight=None):
    """low fp32 [N,h,w] logits, gt fp32 [N,H,W] -> (counts int32 [N,T,3], sums fp64 [N] or None): `seg_counts` and
    `bce_logits_sum` of the bilinear up-sampling of `low` to gt's size, without materialising it."""
    import ctypes
    T = len(thresholds)
    if not 1 <= T <= 8:
        raise ValueError("score_lowres: 1..8 thresholds per call")
    N, h, w = low.shape
    if gt.dim() != 3 or gt.shape[0] != N:
        raise ValueError(f"score_lowres: low {tuple(low.shape)} vs gt {tuple(gt.shape)}")
    H, W = gt.shape[1:]
    counts = torch.empty((N, T, 3), dtype=torch.int32, device=low.device)
    sums = torch.empty((N,), dtype=torch.float64, device=low.device) if pos_weight is not None else None
    thr = (ctypes.c_float * T)(*[float(t) for t in thresholds])
    native.call("ms2_score_lowres", _chk(low, "low", torch.float32), _chk(gt, "gt", torch.float32),
                ctypes.cast(thr, ctypes.c_void_p), T, float(pos_weight if pos_weight is not None else 1.0),
                counts.data_ptr(), None if sums is None else sums.data_ptr(), N, h, w, H, W, _st())
    return counts, sums


def non_overlap(pred_masks):
    """fp32 [n_obj, ...] scores -> same shape: per pixel the first arg-max object keeps its score, the others are
    clamped to <= -10 (SAM2Base._apply_non_overlapping_constraints)."""
    n = pred_masks.shape[0]
    out = torch.empty_like(pred_masks)
    native.call("ms2_non_overlap", _chk(pred_masks, "pred_masks", torch.float32), out.data_ptr(), n,
                pred_masks.numel() // max(n, 1), _st())
    return out


def mask_stats(x, thr, off):
    """x fp32 logits [N,H,W] -> int32 [N,7] = (#(x>thr+off), #(x>thr-off), #(x>thr), min col, min row, max col, max row
    of x>thr); an empty plane gives (.., 0, W, H, -1, -1)."""
    N, H, W = x.shape
    stats = torch.empty((N, 7), dtype=torch.int32, device=x.device)
    native.call("ms2_mask_stats", _chk(x, "x", torch.float32), stats.data_ptr(), N, H, W, float(thr), float(off), _st())
    return stats


def mask_binarize_t(x, sel, thr, out_hw, origin):
    """x fp32 logits [N,H,W], sel int32 [K] plane indices -> uint8 [K,OW,OH]: (x[sel] > thr) placed at origin = (x0, y0)
    of a zero [OH,OW] canvas and transposed (column-major order of the canvas, the order RLE encodes in)."""
    N, H, W = x.shape
    OH, OW = int(out_hw[0]), int(out_hw[1])
    K = sel.numel()
    out = torch.empty((K, OW, OH), dtype=torch.uint8, device=x.device)
    native.call("ms2_mask_binarize_t", _chk(x, "x", torch.float32), _chk(sel, "sel", torch.int32), out.data_ptr(), K, H, W,
                float(thr), OH, OW, int(origin[0]), int(origin[1]), _st())
    return out


def rle_transitions(m, cap):
    """m uint8 [K,L] (0/1, already in encoding order) -> (pos int32 [K,cap], cnt int32 [K]): sorted positions p >= 1 with
    m[k,p] != m[k,p-1]; cnt is the total per mask (cnt[k] > cap: the tail was dropped, call again with a larger cap)."""
    K, L = m.shape
    pos = torch.empty((K, cap), dtype=torch.int32, device=m.device)
    cnt = torch.empty((K,), dtype=torch.int32, device=m.device)
    native.call("ms2_rle_transitions", _chk(m, "m", torch.uint8), pos.data_ptr(), cnt.data_ptr(), K, L, int(cap), _st())
    return pos, cnt
