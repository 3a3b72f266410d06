"""Memory attention stack (reference modeling/memory_attention.py:15-169), token-major.

Per layer: LN -> fused-qkv GEMM -> in-place RoPE -> flash self-attention (D=256, 1 head) -> out-proj
with the residual in the GEMM epilogue; LN -> q GEMM+RoPE, K/V projections of the memory bank
(keys carry the spatial+temporal position code, RoPE restarts every 4096 rows, the trailing
object-pointer tokens are not rotated) -> flash cross-attention -> out-proj+residual; LN -> FFN
with ReLU and residual fused in the two GEMM epilogues.
"""
import torch
from torch import nn

from .. import ops
from ..runtime import compute_dtype
from .sam.transformer import RoPEAttention
from .sam2_utils import LayerNorm, Linear, seq_to_tokens, to_compute


class MemoryAttentionLayer(nn.Module):
    def __init__(self, activation, cross_attention, d_model, dim_feedforward, dropout, pos_enc_at_attn,
                 pos_enc_at_cross_attn_keys, pos_enc_at_cross_attn_queries, self_attention):
        super().__init__()
        assert activation == "relu", "only the shipped (relu) FFN is built"
        self.d_model, self.dim_feedforward = d_model, dim_feedforward
        self.self_attn = self_attention
        self.cross_attn_image = cross_attention
        self.linear1 = Linear(d_model, dim_feedforward)
        self.linear2 = Linear(dim_feedforward, d_model)
        self.norm1, self.norm2, self.norm3 = LayerNorm(d_model), LayerNorm(d_model), LayerNorm(d_model)
        self.pos_enc_at_attn = pos_enc_at_attn
        self.pos_enc_at_cross_attn_queries = pos_enc_at_cross_attn_queries
        self.pos_enc_at_cross_attn_keys = pos_enc_at_cross_attn_keys

    def forward_tokens(self, x, query_pos, mem_k_in, mem_v_in, num_k_exclude_rope):
        """x fp32 [B,L,C]; mem_k_in / mem_v_in compute-dtype [B,Lk,64]."""
        cd = compute_dtype()
        L = x.shape[1]
        t = self.norm1(x, out_dtype=cd)
        if self.pos_enc_at_attn:
            t = ops.axpby(self.norm1(x), 1.0, query_pos, 1.0, out_dtype=cd)
            x = self.self_attn(t, t, self.norm1(x, out_dtype=cd), residual=x)
        else:
            x = self.self_attn.forward_self(t, residual=x)
        t = self.norm2(x, out_dtype=cd)
        if self.pos_enc_at_cross_attn_queries:
            t = ops.axpby(self.norm2(x), 1.0, query_pos, 1.0, out_dtype=cd)
        q = self.cross_attn_image.project_q(t)
        k, v = self.cross_attn_image.project_kv(mem_k_in, mem_v_in, L, num_k_exclude_rope)
        o = ops.attention(q, k, v, self.cross_attn_image.num_heads)
        x = self.cross_attn_image.out_proj(o, out_dtype=torch.float32, residual=x)
        t = self.norm3(x, out_dtype=cd)
        h = self.linear1(t, out_dtype=cd, act=ops.ACT_RELU)
        return self.linear2(h, out_dtype=torch.float32, residual=x)


class MemoryAttention(nn.Module):
    def __init__(self, d_model, pos_enc_at_input, layer, num_layers, batch_first=True):
        super().__init__()
        import copy
        self.d_model = d_model
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = LayerNorm(d_model)
        self.pos_enc_at_input = pos_enc_at_input
        self.batch_first = batch_first

    def forward_tokens(self, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens=0):
        """batch-first fp32: curr/curr_pos [B,L,C], memory/memory_pos [B,Lk,Cm] -> fp32 [B,L,C]."""
        cd = compute_dtype()
        x = ops.axpby(curr, 1.0, curr_pos, 0.1) if (self.pos_enc_at_input and curr_pos is not None) else curr
        keys_at_pos = self.layers[0].pos_enc_at_cross_attn_keys
        mem_v_in = to_compute(memory)
        mem_k_in = ops.axpby(memory, 1.0, memory_pos, 1.0, out_dtype=cd) if keys_at_pos else mem_v_in
        for layer in self.layers:
            x = layer.forward_tokens(x, curr_pos, mem_k_in, mem_v_in, num_obj_ptr_tokens)
        return self.norm(x)

    def forward(self, curr, memory, curr_pos=None, memory_pos=None, num_obj_ptr_tokens=0):
        """Reference signature: sequence-first [L,B,C] tensors (or 1-element lists of them)."""
        if isinstance(curr, list):
            assert isinstance(curr_pos, list) and len(curr) == len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        assert curr.shape[1] == memory.shape[1], "Batch size must be the same for curr and memory"
        out = self.forward_tokens(seq_to_tokens(curr.float()), seq_to_tokens(curr_pos.float()),
                                  seq_to_tokens(memory.float()), seq_to_tokens(memory_pos.float()), num_obj_ptr_tokens)
        return out.permute(1, 0, 2)
