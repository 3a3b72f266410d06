"""Plain-torch statements of every op in `medsam2_b200.ops` (same signatures).

TEST INFRASTRUCTURE ONLY.  Two uses:
  * `-m gpu`: each CUDA kernel is compared with its function here on the same seeded inputs;
  * `-m "not gpu"`: `install(monkeypatch)` swaps these in for the native ops so that the HOST logic
    of the product (layout plumbing, state machine, checkpoint handling) is checked end-to-end against
    the oracle on CPU.  The product package itself never imports this file and has no CPU path.
"""
import math

import torch
import torch.nn.functional as F

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3


def _act(x, act):
    if act == ACT_GELU:
        return F.gelu(x)
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def cc_label(mask_u8):
    from oracle.sam2_oracle import connected_components_np
    l, c = connected_components_np(mask_u8.cpu().numpy())
    return torch.from_numpy(l).to(mask_u8.device), torch.from_numpy(c).to(mask_u8.device)


def fill_holes(scores, max_area, thresh=0.0, fill_value=0.1):
    lab, area = cc_label((scores <= thresh).to(torch.uint8))
    hole = (lab > 0) & (area <= max_area)
    return torch.where(hole, torch.full_like(scores, fill_value), scores)


def layernorm(x, gamma, beta, eps, out_dtype=torch.float32, add=None, act=ACT_NONE):
    h = x.float() if add is None else x.float() + add.float()
    y = F.layer_norm(h, (h.shape[-1],), gamma.float(), beta.float(), eps)
    return _act(y, act).to(out_dtype)


def gemm(a, w, bias=None, out_dtype=torch.float32, act=ACT_NONE, residual=None, colscale=None, impl=0, out=None):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    y = _act(y, act)
    if colscale is not None:
        y = y * colscale
    if residual is not None:
        y = y + residual.reshape(y.shape)
    y = y.to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def attention(q, k, v, heads, scale=None, impl=0):
    B, Lq, HD = q.shape
    D = HD // heads
    sp = lambda t: t.float().reshape(B, t.shape[1], heads, D).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v), scale=scale)
    return o.transpose(1, 2).reshape(B, Lq, HD).to(q.dtype)


def window_attention(qkv, qkv_bias, B, H, W, heads, D, ws, qpool):
    """Restates hieradet.py:58-83,136-159 + backbones/utils.py:16-62 on the qkv tensor: zero-padded
    tokens are replaced by the qkv bias (= Linear(0))."""
    C3 = 3 * heads * D
    x = qkv.float().reshape(B, H, W, C3)
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    Hp, Wp = H + ph, W + pw
    xp = qkv_bias.float().reshape(1, 1, 1, C3).expand(B, Hp, Wp, C3).clone()
    xp[:, :H, :W] = x
    xw = xp.view(B, Hp // ws, ws, Wp // ws, ws, C3).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, 3, heads, D)
    q, k, v = xw.unbind(2)
    w_out = ws
    if qpool:
        nW = q.shape[0]
        q = q.reshape(nW, ws, ws, heads * D).permute(0, 3, 1, 2)
        q = F.max_pool2d(q, 2, 2).permute(0, 2, 3, 1)
        w_out = ws // 2
        q = q.reshape(nW, w_out * w_out, heads, D)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    o = o.transpose(1, 2).reshape(-1, w_out, w_out, heads * D)
    Ho, Wo = (H // 2, W // 2) if qpool else (H, W)
    nwy, nwx = Hp // ws, Wp // ws
    o = o.view(B, nwy, nwx, w_out, w_out, heads * D).permute(0, 1, 3, 2, 4, 5).reshape(B, nwy * w_out, nwx * w_out, -1)
    return o[:, :Ho, :Wo].contiguous().to(qkv.dtype)


def maxpool2x2(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def patch_embed(img, w, bias, pos):
    y = F.conv2d(img, w, bias, stride=4, padding=3).permute(0, 2, 3, 1)
    if pos is not None:
        y = y + pos[None]
    return y.contiguous()


def axpby(x, a=1.0, z=None, b=1.0, c=0.0, out_dtype=torch.float32, out=None):
    y = a * x.float()
    if z is not None:
        zz = z.float().reshape(-1)
        y = (y.reshape(-1, zz.numel()) + b * zz).reshape(x.shape)
    y = (y + c).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gate_rows(x, gate, fill):
    g = gate.reshape(x.shape[0], *([1] * (x.dim() - 1)))
    return torch.where(g > 0, x, torch.full_like(x, fill))


def select_plane(x, idx):
    B = x.shape[0]
    return x[torch.arange(B, device=x.device), idx.long().clamp(0, x.shape[1] - 1)].unsqueeze(1).contiguous()


def add_rowvec(x, v, s=1.0):
    return x + s * v


def cast(x, dtype):
    return x.to(dtype)


def activation(x, act):
    return _act(x, act)


def upsample2x_add_(fine, coarse):
    up = coarse.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    fine.add_(up)
    return fine


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rope_(x, B, rows, n_rope_rows, D, cos_t, sin_t, batch_stride=None, row_stride=None):
    bs = batch_stride if batch_stride is not None else rows * D
    rs = row_stride if row_stride is not None else D
    v = torch.as_strided(x, (B, n_rope_rows, D // 2, 2), (bs, rs, 2, 1), x.storage_offset())
    T = cos_t.shape[0]
    pos = torch.arange(n_rope_rows, device=x.device) % T
    c, s = cos_t[pos][None], sin_t[pos][None]
    a, b = v[..., 0].float(), v[..., 1].float()
    v[..., 0], v[..., 1] = (a * c - b * s).to(x.dtype), (a * s + b * c).to(x.dtype)
    return x


def im2col(x, k, stride, pad, out_dtype, pre=0, pre_scale=1.0, pre_bias=0.0):
    B, H, W, Cin = x.shape
    t = x.float()
    if pre == 1:
        t = torch.sigmoid(t)
    elif pre == 2:
        t = (t > 0).float()
    if pre or pre_scale != 1.0 or pre_bias != 0.0:
        t = t * pre_scale + pre_bias
    u = F.unfold(t.permute(0, 3, 1, 2), k, padding=pad, stride=stride)       # [B, Cin*k*k, L] (ci,ky,kx)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    u = u.view(B, Cin, k * k, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(B, Ho, Wo, k * k * Cin)
    return u.contiguous().to(out_dtype)


def dwconv7x7(x, w, bias):
    C = x.shape[-1]
    y = F.conv2d(x.permute(0, 3, 1, 2), w.view(C, 1, 7, 7), bias, padding=3, groups=C)
    return y.permute(0, 2, 3, 1).contiguous()


def pixel_shuffle_add(g, bias, skip, B, H, W, C, act=ACT_NONE):
    y = g.view(B, H, W, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C)
    if bias is not None:
        y = y + bias
    if skip is not None:
        y = y + skip.reshape(y.shape)
    return _act(y, act).contiguous()


def hyper_mask(up, hyper):
    return torch.einsum("bmc,bpc->bmp", hyper, up).contiguous()


def resize_bilinear(x, size, antialias=False):
    lead = x.shape[:-2]
    y = F.interpolate(x.reshape(-1, 1, *x.shape[-2:]).float(), size=tuple(size), mode="bilinear",
                      align_corners=False, antialias=bool(antialias))
    return y.reshape(*lead, *size)


def fourier_pe(coords01, gauss):
    c = (2 * coords01 - 1) @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def normalize_image(x, out=None, nhwc=None, out_dtype=torch.float32):
    if x.dtype == torch.uint8:
        if nhwc is None:
            nhwc = x.shape[-1] == 3 and x.shape[1] != 3
        x = x.permute(0, 3, 1, 2).float() if nhwc else x.float()
    mean = torch.tensor((0.485, 0.456, 0.406), device=x.device)[None, :, None, None]
    std = torch.tensor((0.229, 0.224, 0.225), device=x.device)[None, :, None, None]
    y = ((x.float() / 255.0 - mean) / std).to(out_dtype if out is None else out.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def mask_stability_counts(x, delta):
    f = x.reshape(x.shape[0], -1)
    return torch.stack([(f > delta).sum(-1), (f > -delta).sum(-1)], dim=1).to(torch.int32)


def mask_stats(x, thr, off):
    N, H, W = x.shape
    out = torch.zeros((N, 7), dtype=torch.int32)
    for n in range(N):
        m = x[n] > thr
        out[n, 0], out[n, 1], out[n, 2] = int((x[n] > thr + off).sum()), int((x[n] > thr - off).sum()), int(m.sum())
        ys, xs = torch.nonzero(m, as_tuple=True)
        out[n, 3:] = torch.tensor([xs.min(), ys.min(), xs.max(), ys.max()] if len(xs) else [W, H, -1, -1])
    return out.to(x.device)


def mask_binarize_t(x, sel, thr, out_hw, origin):
    N, H, W = x.shape
    OH, OW = out_hw
    out = torch.zeros((sel.numel(), OH, OW), dtype=torch.uint8, device=x.device)
    out[:, origin[1]:origin[1] + H, origin[0]:origin[0] + W] = (x[sel.long()] > thr).to(torch.uint8)
    return out.transpose(1, 2).contiguous()


def rle_transitions(m, cap):
    K, L = m.shape
    pos = torch.zeros((K, cap), dtype=torch.int32, device=m.device)
    cnt = torch.zeros((K,), dtype=torch.int32, device=m.device)
    for k in range(K):
        p = torch.nonzero(m[k, 1:] != m[k, :-1]).flatten() + 1
        cnt[k] = len(p)
        pos[k, :min(cap, len(p))] = p[:cap].to(torch.int32)
    return pos, cnt


def seg_counts(pred, gt, thresholds):
    N = pred.shape[0]
    p, g = pred.reshape(N, -1).float(), gt.reshape(N, -1).float()
    out = torch.zeros((N, len(thresholds), 3), dtype=torch.int32, device=pred.device)
    for t, th in enumerate(thresholds):
        a, b = p > float(th), g > float(th)
        out[:, t, 0], out[:, t, 1], out[:, t, 2] = (a & b).sum(1), a.sum(1), b.sum(1)
    return out


def bce_logits_sum(pred, gt, pos_weight):
    N = pred.shape[0]
    x, y = pred.reshape(N, -1).double(), gt.reshape(N, -1).double()
    el = (1 - y) * x + (1 + (pos_weight - 1) * y) * (torch.log1p(torch.exp(-x.abs())) + torch.clamp_min(-x, 0))
    return el.sum(1)


def score_lowres(low, gt, thresholds, pos_weight=None):
    up = F.interpolate(low[:, None].float(), size=gt.shape[-2:], mode="bilinear", align_corners=False)[:, 0]
    counts = seg_counts(up.reshape(up.shape[0], -1), gt.reshape(gt.shape[0], -1), thresholds)
    sums = None if pos_weight is None else bce_logits_sum(up.reshape(up.shape[0], -1), gt.reshape(gt.shape[0], -1), pos_weight)
    return counts, sums


def non_overlap(pred_masks):
    n = pred_masks.shape[0]
    top = torch.argmax(pred_masks, dim=0, keepdim=True)
    keep = top == torch.arange(n, device=pred_masks.device).view(n, *([1] * (pred_masks.dim() - 1)))
    return torch.where(keep, pred_masks, torch.clamp(pred_masks, max=-10.0))


def bank_rows(srcs, poss, out_dtype, k_out=None, m_out=None):
    ks = []
    for t, q in zip(srcs, poss):
        ks.append(t if q is None else t + (q if q.dim() == 3 else q[None]))
    k = torch.cat(ks, dim=1).to(out_dtype)
    m = torch.cat(list(srcs), dim=1).to(out_dtype)
    if k_out is None:
        k_out = k
    else:
        k_out.copy_(k)
    if m_out is not None:
        m_out.copy_(m)
    return k_out, m_out


def argmax_select_rows(scores, rows=None):
    idx = torch.argmax(scores, dim=-1)
    out = None if rows is None else rows[torch.arange(scores.shape[0], device=scores.device), idx].contiguous()
    return idx.to(torch.int32), out


def obj_ptr_mix(ptr, logits, no_obj, soft, fixed):
    lam = (torch.sigmoid(logits) if soft else (logits > 0).float()).reshape(-1, 1)
    return (lam * ptr if fixed else ptr) + (1 - lam) * no_obj.reshape(1, -1)


def stability_select(counts, ious, thresh):
    c = counts.float()
    stab = torch.where(c[:, 1] > 0, c[:, 0] / c[:, 1], torch.ones_like(c[:, 1]))
    best = torch.argmax(ious[:, 1:], dim=-1) + 1
    idx = torch.where(stab >= thresh, torch.zeros_like(best), best)
    return idx.to(torch.int32), torch.gather(ious, 1, idx[:, None])


def conv3x3s2_ln_gelu(x, w, bias, gamma, beta, eps, out_dtype=torch.float32, pre=0, pre_scale=1.0, pre_bias=0.0):
    xi = x.float()
    if pre == 1:
        xi = torch.sigmoid(xi)
    elif pre == 2:
        xi = (xi > 0).float()
    if pre:
        xi = xi * pre_scale + pre_bias
    y = F.conv2d(xi.permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    y = F.layer_norm(y, (y.shape[-1],), gamma.float(), beta.float(), eps)
    return F.gelu(y).to(out_dtype).contiguous()


def gemm_grouped(a_list, w_list, bias_list, out_dtype=torch.float32, acts=None, outs=None):
    acts = acts or [ACT_NONE] * len(a_list)
    res = [gemm(a, w, b, out_dtype=out_dtype, act=act) for a, w, b, act in zip(a_list, w_list, bias_list, acts)]
    if outs is not None:
        for i, o in enumerate(outs):
            if o is not None:
                o.copy_(res[i])
                res[i] = o
    return res


def attention_dv(q, k, v, scale=None):
    o = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None], scale=scale)[:, 0]
    return o.to(q.dtype)


def attention_dv_partial(q, k, v, out=None, scale=None):
    """plain statement of ms2_attention_dv_partial: packed (O un-normalised, row max in log2 units, row sum)."""
    B, Lq, D = q.shape
    DV, rows = 64, B * Lq
    if out is None:
        out = torch.empty(rows * (DV + 2), dtype=torch.float32, device=q.device)
    po, pml = out[: rows * DV].view(rows, DV), out[rows * DV:].view(rows, 2)
    if k is None or k.shape[1] == 0:
        po.zero_()
        pml[:, 0] = float("-inf")
        pml[:, 1] = 0.0
        return out
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    s = torch.einsum("bqd,bkd->bqk", q.float(), k.float()) * (scale * 1.4426950408889634)
    m = s.amax(-1, keepdim=True)
    p = torch.exp2(s - m)
    po.copy_(torch.einsum("bqk,bkd->bqd", p, v.float()).reshape(rows, DV))
    pml[:, 0] = m.reshape(rows)
    pml[:, 1] = p.sum(-1).reshape(rows)
    return out


def attention_merge(parts, B, Lq, DV=64):
    rows = B * Lq
    o = parts[:, : rows * DV].view(-1, rows, DV)
    ml = parts[:, rows * DV:].view(-1, rows, 2)
    m = ml[..., 0].amax(0)
    w = torch.where(ml[..., 1] > 0, torch.exp2(ml[..., 0] - m), torch.zeros_like(m))
    num = (w[..., None] * o).sum(0)
    den = (w * ml[..., 1]).sum(0)
    return (num / den[:, None]).view(B, Lq, DV).to(torch.bfloat16)


def point_embed(coords, labels, gauss, table, pad, image_size, prefix=None):
    pts = coords.float() + 0.5
    lab = labels.long()
    if pad:
        pts = torch.cat([pts, torch.zeros((pts.shape[0], 1, 2), device=pts.device)], dim=1)
        lab = torch.cat([lab, -torch.ones((lab.shape[0], 1), device=lab.device, dtype=lab.dtype)], dim=1)
    c = pts.clone()
    c[:, :, 0] = c[:, :, 0] / image_size[1]
    c[:, :, 1] = c[:, :, 1] / image_size[0]
    pe = fourier_pe(c.contiguous(), gauss)
    keep = (lab != -1).to(pe.dtype).unsqueeze(-1)
    out = pe * keep + table[(lab + 1).clamp(0, 4)]
    if prefix is not None:
        out = torch.cat([prefix[None].expand(out.shape[0], -1, -1), out], dim=1).contiguous()
    return out


def patch_im2col(img, ldk=152):
    B, _, H, W = img.shape
    cols = F.unfold(img.float(), 7, stride=4, padding=3)                      # [B, 3*49, L] in (c, ky, kx) order
    L = cols.shape[-1]
    cols = cols.view(B, 3, 49, L).permute(0, 3, 2, 1).reshape(B, L, 147)       # -> (ky, kx, c)
    return F.pad(cols, (0, ldk - 147)).to(torch.bfloat16)


def cast_into(x, out):
    out.copy_(x.reshape(out.shape))
    return out


_NAMES = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("install",)]


def install(monkeypatch):
    """Swap the native ops for these statements (CPU host-logic tests only)."""
    import medsam2_b200.ops as ops
    for n in _NAMES:
        if hasattr(ops, n):
            monkeypatch.setattr(ops, n, globals()[n])
