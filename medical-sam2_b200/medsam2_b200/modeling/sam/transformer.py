"""Attention blocks of the mask decoder and the memory-attention stack, token-major.

Parameters mirror the reference (modeling/sam/transformer.py:28-331) so checkpoints load unchanged;
the computation is GEMM (+fused bias/residual) -> optional in-place RoPE -> flash attention kernels.
"""
import math

import torch
from torch import nn

from ... import ops
from ...runtime import cat_p32, cat_w_c, compute_dtype, p32, par, w_c
from ..position_encoding import rope_table
from ..sam2_utils import LayerNorm, Linear, MLP, to_compute


class Attention(nn.Module):
    """q/k/v projections with optional channel down-scaling (transformer.py:199-263)."""

    def __init__(self, embedding_dim, num_heads, downsample_rate=1, dropout=0.0, kv_in_dim=None):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.kv_in_dim = kv_in_dim if kv_in_dim is not None else embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = Linear(embedding_dim, self.internal_dim)
        self.k_proj = Linear(self.kv_in_dim, self.internal_dim)
        self.v_proj = Linear(self.kv_in_dim, self.internal_dim)
        self.out_proj = Linear(self.internal_dim, embedding_dim)
        self.dropout_p = dropout   # inactive in eval; kept for the constructor signature

    def forward(self, q, k, v, residual=None):
        """q [B,Lq,C], k/v [B,Lk,Ckv] (fp32 or compute dtype) -> fp32 [B,Lq,C] (+ residual)."""
        cd = compute_dtype()
        B, Lk = k.shape[0], k.shape[1]
        small_kv = B * Lk <= 64 and k.shape == v.shape and k.shape[-1] % 8 == 0 and v.shape[-1] % 8 == 0
        with_q = small_kv and q.shape == k.shape and self.q_proj.weight.shape[1] == self.k_proj.weight.shape[1]

        def kv_side():
            # token-side projections (a handful of rows: pure launch latency) share one grouped small-M launch
            if not small_kv:
                return None, self.k_proj(k, out_dtype=cd), self.v_proj(v, out_dtype=cd)
            xs = [to_compute(k).reshape(B * Lk, -1), to_compute(v).reshape(B * Lk, -1)]
            ls = [self.k_proj, self.v_proj]
            if with_q:
                xs.insert(0, to_compute(q).reshape(B * Lk, -1))
                ls.insert(0, self.q_proj)
            outs = ops.gemm_grouped(xs, [w_c(l.weight) for l in ls], [p32(l.bias) for l in ls], out_dtype=cd)
            outs = [o.view(B, Lk, -1) for o in outs]
            return (outs[0] if with_q else None), outs[-2], outs[-1]

        if with_q:
            qp, kp, vp = kv_side()
        else:       # queries and keys/values live on different sides (tokens vs image): independent branches
            qp, (_, kp, vp) = par(lambda: self.q_proj(q, out_dtype=cd), kv_side)
        o = ops.attention(qp, kp, vp, self.num_heads)
        return self.out_proj(o, out_dtype=torch.float32, residual=residual)


class RoPEAttention(Attention):
    """Single-head attention with axial RoPE on q and (a prefix of) k (transformer.py:266-331)."""

    def __init__(self, *args, rope_theta=10000.0, rope_k_repeat=False, feat_sizes=(32, 32), **kwargs):
        super().__init__(*args, **kwargs)
        self.rope_theta = rope_theta
        self.rope_k_repeat = rope_k_repeat
        self.feat_sizes = tuple(feat_sizes)

    def _table(self, Lq, device):
        side = int(round(math.sqrt(Lq)))
        assert side * side == Lq, "RoPE attention expects a square token grid"
        return rope_table(self.internal_dim // self.num_heads, side, side, self.rope_theta, device)

    def forward_self(self, t, residual):
        """self-attention: q=k=v-source=t [B,L,C] (compute dtype); one fused qkv GEMM."""
        cd = compute_dtype()
        B, L, C = t.shape
        Ci = self.internal_dim
        w = cat_w_c(self.q_proj.weight, self.k_proj.weight, self.v_proj.weight)
        b = cat_p32(self.q_proj.bias, self.k_proj.bias, self.v_proj.bias)
        qkv = ops.gemm(t, w, b, out_dtype=cd)                       # [B,L,3*Ci]
        cos, sin = self._table(L, t.device)
        D = Ci // self.num_heads
        if B == 1 and self.num_heads == 1:
            # q and k of the fused projection in one launch: "batch" 0 = q columns, 1 = k columns of the same rows
            ops.rope_(qkv, 2, L, L, D, cos, sin, batch_stride=Ci, row_stride=3 * Ci)
        else:
            for part in (0, 1):
                for h in range(self.num_heads):
                    ops.rope_(qkv[:, :, part * Ci + h * D:], B, L, L, D, cos, sin, batch_stride=L * 3 * Ci, row_stride=3 * Ci)
        o = ops.attention(qkv[:, :, :Ci], qkv[:, :, Ci:2 * Ci], qkv[:, :, 2 * Ci:], self.num_heads)
        return self.out_proj(o, out_dtype=torch.float32, residual=residual)

    def project_q(self, t):
        cd = compute_dtype()
        B, L, _ = t.shape
        q = self.q_proj(t, out_dtype=cd)
        cos, sin = self._table(L, t.device)
        D = self.internal_dim // self.num_heads
        for h in range(self.num_heads):
            ops.rope_(q[:, :, h * D:], B, L, L, D, cos, sin, batch_stride=L * self.internal_dim, row_stride=self.internal_dim)
        return q

    def project_kv(self, k_in, v_in, Lq, num_k_exclude_rope=0):
        """k_in/v_in [B,Lk,kv_in_dim] (compute dtype) -> roped K and V [B,Lk,Ci]."""
        cd = compute_dtype()
        B, Lk, _ = k_in.shape
        k = self.k_proj(k_in, out_dtype=cd)
        v = self.v_proj(v_in, out_dtype=cd)
        n_rope = Lk - num_k_exclude_rope
        if n_rope > 0:
            if Lk != Lq:
                assert self.rope_k_repeat
            cos, sin = self._table(Lq, k_in.device)
            D = self.internal_dim // self.num_heads
            for h in range(self.num_heads):
                ops.rope_(k[:, :, h * D:], B, Lk, n_rope, D, cos, sin, batch_stride=Lk * self.internal_dim,
                          row_stride=self.internal_dim)
        return k, v

    def forward(self, q, k, v, num_k_exclude_rope=0, residual=None):
        """Reference signature (batch-first [B,L,C] tensors)."""
        qp = self.project_q(to_compute(q))
        kp, vp = self.project_kv(to_compute(k), to_compute(v), q.shape[1], num_k_exclude_rope)
        o = ops.attention(qp, kp, vp, self.num_heads)
        return self.out_proj(o, out_dtype=torch.float32, residual=residual)


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, embedding_dim, num_heads, mlp_dim=2048, activation=None, attention_downsample_rate=2,
                 skip_first_layer_pe=False):
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = LayerNorm(embedding_dim)
        self.mlp = MLP(embedding_dim, mlp_dim, embedding_dim, num_layers=2, activation="relu")
        self.norm3 = LayerNorm(embedding_dim)
        self.norm4 = LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        """fp32 token-major tensors: queries/query_pe [B,Nt,C]; keys/key_pe [B,HW,C] (transformer.py:165-196)."""
        cd = compute_dtype()
        if self.skip_first_layer_pe:
            qc = to_compute(queries)
            queries = self.norm1(self.self_attn(qc, qc, qc))
        else:
            q = ops.axpby(queries, 1.0, query_pe, 1.0, out_dtype=cd)
            queries = self.norm1(queries, add=self.self_attn(q, q, to_compute(queries)))
        q, k, keys_c = par(lambda: ops.axpby(queries, 1.0, query_pe, 1.0, out_dtype=cd),
                           lambda: ops.axpby(keys, 1.0, key_pe, 1.0, out_dtype=cd), lambda: to_compute(keys))
        queries = self.norm2(queries, add=self.cross_attn_token_to_image(q, k, keys_c))
        queries = self.norm3(queries, add=self.mlp(queries))
        q = ops.axpby(queries, 1.0, query_pe, 1.0, out_dtype=cd)
        keys = self.norm4(keys, add=self.cross_attn_image_to_token(k, q, to_compute(queries)))
        return queries, keys


class TwoWayTransformer(nn.Module):
    def __init__(self, depth, embedding_dim, num_heads, mlp_dim, activation=None, attention_downsample_rate=2):
        super().__init__()
        self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = depth, embedding_dim, num_heads, mlp_dim
        self.layers = nn.ModuleList(
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, attention_downsample_rate=attention_downsample_rate,
                                 skip_first_layer_pe=(i == 0)) for i in range(depth))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = LayerNorm(embedding_dim)

    def forward_tokens(self, keys, key_pe, point_embedding):
        """keys/key_pe fp32 [B,HW,C]; point_embedding fp32 [B,Nt,C] -> (queries, keys)."""
        cd = compute_dtype()
        queries = point_embedding
        for layer in self.layers:
            queries, keys = layer(queries, keys, point_embedding, key_pe)
        q, k, keys_c = par(lambda: ops.axpby(queries, 1.0, point_embedding, 1.0, out_dtype=cd),
                           lambda: ops.axpby(keys, 1.0, key_pe, 1.0, out_dtype=cd), lambda: to_compute(keys))
        queries = self.norm_final_attn(queries, add=self.final_attn_token_to_image(q, k, keys_c))
        return queries, keys

    def forward(self, image_embedding, image_pe, point_embedding):
        """Reference signature: NCHW image tensors (transformer.py:74-118)."""
        from ..sam2_utils import as_nhwc
        B, C, H, W = image_embedding.shape
        keys = as_nhwc(image_embedding).reshape(B, H * W, C)
        pe = as_nhwc(image_pe.expand(B, -1, -1, -1)).reshape(B, H * W, C)
        return self.forward_tokens(keys, pe, point_embedding.contiguous())
