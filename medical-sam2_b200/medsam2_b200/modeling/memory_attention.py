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
from ..runtime import compute_dtype, p32
from .sam.transformer import RoPEAttention
from .sam2_utils import LayerNorm, Linear, seq_to_tokens, to_compute


import os as _os
_NO_NATIVE = _os.environ.get("MS2_MEMATTN_NATIVE", "1") == "0"      # 0: drive the stack launch by launch from Python


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

    def forward_tokens(self, x, query_pos, mem_k_in, mem_v_in, num_k_exclude_rope, kv=None):
        """x fp32 [B,L,C]; mem_k_in / mem_v_in compute-dtype [B,Lk,64], or kv = (K, V) already projected
        (and rotated) views of a MemoryBank: (K [B,Lk,C], V [B,Lk,C], False) or (K, raw memory [B,Lk,Cm], True)."""
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
        if kv is not None and kv[2]:
            attend = kv[3] if len(kv) > 3 else ops.attention_dv          # kv[3]: split-KV over several GPUs (KVShard.attend)
            # deferred value projection and output projection are two linear maps in a row on the [L,64] attention
            # result: ONE GEMM with W = Wo Wv [256,64], b = Wo bv + bo (folded in fp32), residual in the epilogue
            att = self.cross_attn_image

            def fold(wv, bv, wo, bo):
                Wo, Wv = wo.detach().float(), wv.detach().float()
                return (Wo @ Wv).to(cd).contiguous(), (Wo @ bv.detach().float() + bo.detach().float()).contiguous()
            from ..runtime import CACHE
            W, b = CACHE.get((att.v_proj.weight, att.v_proj.bias, att.out_proj.weight, att.out_proj.bias), ("vo_fold", cd), fold)
            x = ops.gemm(attend(q, kv[0], kv[1]), W, b, out_dtype=torch.float32, residual=x)
        else:
            k, v = kv[:2] if kv is not None else self.cross_attn_image.project_kv(mem_k_in, mem_v_in, L, num_k_exclude_rope)
            o = ops.attention(q, k, v, self.cross_attn_image.num_heads)
            x = self.cross_attn_image.out_proj(o, out_dtype=torch.float32, residual=x)
        t = self.norm3(x, out_dtype=cd)
        h = self.linear1(t, out_dtype=cd, act=ops.ACT_RELU)
        return self.linear2(h, out_dtype=torch.float32, residual=x)


class MemoryBank:
    """Resident per-layer K (RoPE applied) / V of the memory bank of one tracking session.

    Rows [0, n_static) hold the conditioning memories in arrival order; they are projected ONCE, because for a
    conditioning frame neither the temporal slot (t_pos = 0), nor the spatial code, nor the RoPE phase (restarts
    every frame) depends on the frame being tracked (reference sam2_base.py:566-580, memory_attention.py:73-79,
    transformer.py:309-315).  Rows [n_static, n_static + n_dyn) are rewritten every frame with the
    <= num_maskmem-1 recent memories and the object-pointer tokens.  Softmax is invariant to the key order, so this
    layout is the reference's bank up to a permutation."""

    def __init__(self):
        self.keys, self.refs = [], []
        self.K, self.V = [], []
        self.M = None                # raw memory values [B,cap,Cm] (compute dtype) when the value projection is deferred
        self.raw_v = False
        self.cap = self.n_static = 0
        self.B = self.device = self.dtype = None

    def reset(self):
        self.__init__()

    def ensure(self, n_layers, B, rows, width, mem_width, raw_v, dtype, device):
        if (self.B, self.device, self.dtype, self.raw_v) != (B, device, dtype, raw_v) or len(self.K) != n_layers:
            self.reset()
            self.B, self.device, self.dtype, self.raw_v = B, device, dtype, raw_v
            self.K, self.V = [None] * n_layers, [None] * n_layers
        if rows > self.cap:
            cap = max(rows, int(self.cap * 1.5))

            def grow(old, w):
                new = torch.empty((B, cap, w), dtype=dtype, device=device)
                if old is not None and self.n_static:
                    new[:, : self.n_static] = old[:, : self.n_static]
                return new
            for l in range(n_layers):
                self.K[l] = grow(self.K[l], width)
                if not raw_v:
                    self.V[l] = grow(self.V[l], width)
            if raw_v:
                self.M = grow(self.M, mem_width)
            self.cap = cap


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
        self.kv_shard = None          # parallel.KVShard: the bank's keys are dealt to several GPUs (split-KV attention)

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

    # ------------------------------------------------------------------ native orchestration (csrc/memattn.cu)
    def _native_ok(self, L):
        """the shipped configuration, for which the stack is driven by ONE C-ABI call per frame (ms2_memattn_forward)"""
        l0 = self.layers[0]
        a, c = l0.self_attn, l0.cross_attn_image
        side = int(round(L ** 0.5))
        return (compute_dtype() == torch.bfloat16 and not self.training and side * side == L and L >= 64
                and a.num_heads == 1 and c.num_heads == 1 and a.internal_dim == self.d_model == c.internal_dim
                and c.kv_in_dim * 4 == self.d_model and self.pos_enc_at_input
                and not l0.pos_enc_at_attn and not l0.pos_enc_at_cross_attn_queries and l0.pos_enc_at_cross_attn_keys)

    def _native_layers(self, L, device):
        """ctypes array of ms2_memattn_layer_w (one record per layer) over the kernel-ready parameter copies; rebuilt when
        the parameter generation changes (the record holds raw pointers, the copies are kept alive next to it)."""
        import ctypes
        from ..runtime import CACHE, cat_p32, cat_w_c, generation, p32, w_c
        key = (generation(), L, str(device))
        hit = getattr(self, "_native_rec", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        cd = compute_dtype()

        class LayerW(ctypes.Structure):
            _fields_ = ([(n, ctypes.c_void_p) for n in (
                "norm1_g", "norm1_b", "norm2_g", "norm2_b", "norm3_g", "norm3_b", "qkv_w", "qkv_b", "so_w", "so_b", "cq_w",
                "cq_b", "ck_w", "ck_b", "vo_w", "vo_b", "f1_w", "f1_b", "f2_w", "f2_b", "rope_cos", "rope_sin")]
                + [(n, ctypes.c_float) for n in ("eps1", "eps2", "eps3")]
                + [(n, ctypes.c_int) for n in ("C", "Cm", "F", "rope_len")])
        arr = (LayerW * len(self.layers))()
        keep = []
        for i, layer in enumerate(self.layers):
            sa, ca = layer.self_attn, layer.cross_attn_image

            def fold(wv, bv, wo, bo):
                Wo, Wv = wo.detach().float(), wv.detach().float()
                return (Wo @ Wv).to(cd).contiguous(), (Wo @ bv.detach().float() + bo.detach().float()).contiguous()
            vo_w, vo_b = CACHE.get((ca.v_proj.weight, ca.v_proj.bias, ca.out_proj.weight, ca.out_proj.bias), ("vo_fold", cd), fold)
            cos, sin = ca._table(L, device)
            t = dict(
                norm1_g=p32(layer.norm1.weight), norm1_b=p32(layer.norm1.bias), norm2_g=p32(layer.norm2.weight),
                norm2_b=p32(layer.norm2.bias), norm3_g=p32(layer.norm3.weight), norm3_b=p32(layer.norm3.bias),
                qkv_w=cat_w_c(sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight),
                qkv_b=cat_p32(sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias),
                so_w=w_c(sa.out_proj.weight), so_b=p32(sa.out_proj.bias), cq_w=w_c(ca.q_proj.weight), cq_b=p32(ca.q_proj.bias),
                ck_w=w_c(ca.k_proj.weight), ck_b=p32(ca.k_proj.bias), vo_w=vo_w, vo_b=vo_b,
                f1_w=w_c(layer.linear1.weight), f1_b=p32(layer.linear1.bias), f2_w=w_c(layer.linear2.weight),
                f2_b=p32(layer.linear2.bias), rope_cos=cos, rope_sin=sin)
            for n, v in t.items():
                assert v.is_cuda and v.is_contiguous(), n
                setattr(arr[i], n, v.data_ptr())
            keep.append(t)
            arr[i].eps1, arr[i].eps2, arr[i].eps3 = layer.norm1.eps, layer.norm2.eps, layer.norm3.eps
            arr[i].C, arr[i].Cm, arr[i].F, arr[i].rope_len = self.d_model, ca.kv_in_dim, layer.dim_feedforward, cos.shape[0]
        self._native_rec = (key, (arr, keep, ctypes.sizeof(LayerW)))
        return self._native_rec[1]

    def _native_ws(self, B, L, n_dyn_rows, device):
        """persistent workspace of the native calls (stream-ordered reuse; memory attention runs on one stream)"""
        from .. import native
        l0 = self.layers[0]
        need = int(native.lib().ms2_memattn_workspace_bytes(B, L, self.d_model, l0.cross_attn_image.kv_in_dim, l0.dim_feedforward))
        need = max(need, B * n_dyn_rows * l0.cross_attn_image.kv_in_dim * 2 + 4096)
        ws = getattr(self, "_native_wsbuf", None)
        if ws is None or ws.numel() < need or ws.device != device:
            ws = self._native_wsbuf = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    def _project_into_bank(self, bank, row0, srcs, poss, n_rope_rows, Lq):
        """srcs: fp32 [B,rows_i,Cm] memories (in key order), poss: per source a position table [rows_i,Cm] / [B,rows_i,Cm]
        or None -> per layer K (RoPE on the first n_rope_rows rows, restarting every Lq rows) and V written at bank rows
        [row0, row0+n).  The concatenation, `+ pos`, and the casts are ONE kernel (ms2_bank_rows)."""
        cd = compute_dtype()
        B = srcs[0].shape[0]
        n = sum(t.shape[1] for t in srcs)
        if bank.raw_v:
            k_in, _ = ops.bank_rows(srcs, poss, cd, m_out=bank.M[:, row0: row0 + n])
            v_in = None
        else:
            v_in = torch.empty((B, n, srcs[0].shape[2]), dtype=cd, device=srcs[0].device)
            k_in, _ = ops.bank_rows(srcs, poss, cd, m_out=v_in)
        for l, layer in enumerate(self.layers):
            att = layer.cross_attn_image
            D = att.internal_dim // att.num_heads
            cos, sin = att._table(Lq, srcs[0].device)
            for b in range(B):
                kd = bank.K[l][b, row0: row0 + n]
                att.k_proj(k_in[b], out_dtype=cd, out=kd)
                if not bank.raw_v:
                    att.v_proj(v_in[b], out_dtype=cd, out=bank.V[l][b, row0: row0 + n])
                if n_rope_rows > 0:
                    for h in range(att.num_heads):
                        ops.rope_(kd[:, h * D:], 1, n, n_rope_rows, D, cos, sin, batch_stride=n * att.internal_dim,
                                  row_stride=att.internal_dim)

    def forward_tokens_banked(self, curr, curr_pos, bank, cond, recent, ptrs, ptr_pos):
        """Same result as forward_tokens on the concatenated bank (up to the key order), with the conditioning
        memories' per-layer K/V taken from (or appended once to) `bank`.  cond / recent: lists of
        (key, source tensor, feats [B,hw,Cm] fp32, pos table [hw,Cm]); ptrs / ptr_pos [B,n_tok,Cm] or None."""
        cd = compute_dtype()
        B, L, C = curr.shape
        keys_at_pos = self.layers[0].pos_enc_at_cross_attn_keys
        att0 = self.layers[0].cross_attn_image
        shard = self.kv_shard
        if shard is not None:
            # this rank's share of the bank: conditioning memories i = rank (mod world), recent memories rotated, the
            # pointer tokens with the first recent memory; softmax partials are merged across ranks per layer
            n_cond, n_recent = len(cond), len(recent)
            cond = [e for i, e in enumerate(cond) if shard.owns_cond(i)]
            recent = [e for j, e in enumerate(recent) if shard.owns_recent(j, n_cond)]
            if not shard.owns_pointers(n_recent, n_cond):
                ptrs = ptr_pos = None
        hw = cond[0][2].shape[1] if cond else (recent[0][2].shape[1] if recent else L)
        ptr_list = ptrs if isinstance(ptrs, list) else ([ptrs] if ptrs is not None else [])
        n_ptr_rows = sum(t.shape[1] for t in ptr_list)
        n_dyn = sum(e[2].shape[1] for e in recent) + n_ptr_rows
        keys = [e[0] for e in cond]
        if bank.keys != keys[: len(bank.keys)] or bank.B not in (None, B):
            bank.reset()
        # bf16 + one head of 256 over 64-d memories (the shipped configuration): keep the raw memory values once
        # instead of a projected V per layer and attend over them (ms2_attention_dv); v_proj then runs on the
        # [L,64] attention result
        raw_v = (cd == torch.bfloat16 and att0.num_heads == 1 and att0.internal_dim == 256 and att0.kv_in_dim == 64
                 and L >= 64)
        if shard is not None and not raw_v:
            raise RuntimeError("split-KV memory attention needs the bf16 single-head 256/64 configuration")
        bank.ensure(len(self.layers), B, len(keys) * hw + n_dyn + 64, att0.internal_dim, att0.kv_in_dim, raw_v, cd,
                    curr.device)
        new = cond[len(bank.keys):]
        if new:                                            # conditioning memories not yet resident: project once
            srcs = [e[2] for e in new]
            poss = [e[3] if keys_at_pos else None for e in new]
            self._project_into_bank(bank, bank.n_static, srcs, poss, sum(t.shape[1] for t in srcs), L)
            bank.n_static += sum(t.shape[1] for t in srcs)
            bank.keys = keys
            bank.refs = [e[1] for e in cond]               # keep the sources alive: id() keys stay unique
        row0 = bank.n_static
        # (per-launch CUDA-event profiling of bench.py needs the tensor-level calls: it switches the native driver off)
        fast = (raw_v and curr.is_cuda and curr_pos is not None and self._native_ok(L) and (ptr_pos is None or isinstance(ptrs, list))
                and ops.PROFILE.names is None and not _NO_NATIVE)
        Lk = row0 + n_dyn
        if fast:
            return self._forward_native(curr, curr_pos, bank, recent, ptr_list, row0, n_dyn, Lk, shard)
        if n_dyn:
            srcs = [e[2] for e in recent] + ptr_list
            if keys_at_pos:
                poss = [e[3] for e in recent] + ([ptr_pos] if (ptr_pos is not None and not isinstance(ptrs, list))
                                                 else [None] * len(ptr_list))
            else:
                poss = [None] * len(srcs)
            self._project_into_bank(bank, row0, srcs, poss, sum(e[2].shape[1] for e in recent), L)
        x = ops.axpby(curr, 1.0, curr_pos, 0.1) if (self.pos_enc_at_input and curr_pos is not None) else curr
        for l, layer in enumerate(self.layers):
            if shard is not None:
                kv = (bank.K[l][:, :Lk] if Lk else None, bank.M[:, :Lk] if Lk else None, True, shard.attend)
            elif bank.raw_v:
                kv = (bank.K[l][:, :Lk], bank.M[:, :Lk], True)
            else:
                kv = (bank.K[l][:, :Lk], bank.V[l][:, :Lk], False)
            x = layer.forward_tokens(x, curr_pos, None, None, 0, kv=kv)
        return self.norm(x)

    def _forward_native(self, curr, curr_pos, bank, recent, ptr_list, row0, n_dyn, Lk, shard):
        """the per-frame part of forward_tokens_banked through csrc/memattn.cu: ONE call projects the frame's dynamic bank
        rows, ONE call runs the four layers + final norm (single GPU); with a KVShard the layers run as pre / exchange /
        post.  Same kernels, same order, same split-KV budgets as the tensor-level path (bit-identical results)."""
        import ctypes
        from .. import native
        B, L, C = curr.shape
        dev = curr.device
        arr, _, rec_size = self._native_layers(L, dev)
        nL = len(self.layers)
        ws = self._native_ws(B, L, n_dyn, dev)
        st = ops._st()
        vp = ctypes.c_void_p
        K_ptrs = (vp * nL)(*[bank.K[l].data_ptr() for l in range(nL)])
        k_bs, m_bs = bank.K[0].stride(0), bank.M.stride(0)
        if n_dyn:
            srcs = [e[2] for e in recent] + ptr_list
            poss = [e[3] for e in recent] + [None] * len(ptr_list)
            n = len(srcs)
            for t in srcs:
                ops._chk(t, "memory", torch.float32)
            native.call("ms2_memattn_bank_project", ctypes.addressof(arr), nL, (vp * n)(*[t.data_ptr() for t in srcs]),
                        (vp * n)(*[None if q is None else ops._chk(q, "pos", torch.float32) for q in poss]),
                        (ctypes.c_long * n)(*([0] * n)), (ctypes.c_int * n)(*[t.shape[1] for t in srcs]), n, B,
                        sum(e[2].shape[1] for e in recent), L, K_ptrs, k_bs, row0, bank.M.data_ptr(), m_bs, ws.data_ptr(),
                        ws.numel(), st)
            native.launch_count += n // 80 + 2 * nL * B              # kernels behind that one call (bank_rows, k_proj, rope)
        x = torch.empty((B, L, C), dtype=torch.float32, device=dev)
        ops._chk(curr, "curr", torch.float32)
        ops._chk(curr_pos, "curr_pos", torch.float32)
        if shard is None:
            out = torch.empty_like(x)
            ev = ops.PROFILE.begin("mem_attention_stack")
            native.call("ms2_memattn_forward", ctypes.addressof(arr), nL, curr.data_ptr(), curr_pos.data_ptr(), 0.1, K_ptrs, k_bs,
                        bank.M.data_ptr(), m_bs, Lk, ops._chk(p32(self.norm.weight), "g"), ops._chk(p32(self.norm.bias), "b"),
                        float(self.norm.eps), x.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), B, L, st)
            ops.PROFILE.end("mem_attention_stack", ev, 0.0)
            native.launch_count += 1 + nL * 16               # axpby, per layer 16 kernels (incl. two split-KV merges), final LN
            return out
        ops.axpby(curr, 1.0, curr_pos, 0.1, out=x)
        q = torch.empty((B, L, C), dtype=torch.bfloat16, device=dev)
        for l in range(nL):
            rec = ctypes.addressof(arr) + l * rec_size
            native.call("ms2_memattn_layer_pre", rec, x.data_ptr(), q.data_ptr(), ws.data_ptr(), ws.numel(), B, L, st)
            att = shard.attend(q, bank.K[l][:, :Lk] if Lk else None, bank.M[:, :Lk] if Lk else None)
            native.call("ms2_memattn_layer_post", rec, x.data_ptr(), ops._chk(att, "att", torch.bfloat16), ws.data_ptr(),
                        ws.numel(), B, L, st)
            native.launch_count += 8 + 3
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
