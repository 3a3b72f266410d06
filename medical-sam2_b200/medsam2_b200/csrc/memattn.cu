// Host-side orchestration of the memory-attention stack (modeling/memory_attention.py:15-169 of the reference) in ONE
// C-ABI call per tracked frame.  Nothing here is a kernel: the functions below launch the library's own kernels
// (ms2_layernorm, ms2_gemm, ms2_rope, ms2_attention_ws, ms2_attention_dv, ms2_bank_rows, ms2_axpby) in the order the
// Python module used to — which cost ~30 us of interpreter time per launch and made the tracked-frame path HOST-bound
// (bench step: 234 ms to enqueue, 0.1 ms of GPU work left when the host was done; profiles/r2_host_profile.txt).
// Shipped configuration only: one head of 256 over 64-d memories, pos_enc_at_input, keys carry the position code,
// no position code at the self-attention / cross-attention queries, ReLU FFN, bf16 operands.
#include "common.cuh"

extern "C" {
int ms2_layernorm(const float* x, const float* add, const float* gamma, const float* beta, void* y, int y_dt, int M, int C,
                  float eps, int act, void* stream);
int ms2_gemm(const void* A, int a_dt, long lda, const void* W, int w_dt, const float* bias, const float* colscale,
             const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K, int act, int impl,
             void* stream);
int ms2_rope(void* x, int dt, long batch_stride, long row_stride, int B, int rows, int n_rope_rows, int D,
             const float* cos_t, const float* sin_t, int table_len, void* stream);
int ms2_attention_ws(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs, long q_ts, long k_bs,
                     long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts, int B, int Hh,
                     int Lq, int Lk, int D, float scale, int impl, void* workspace, long workspace_bytes, void* stream);
int ms2_attention_dv(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs, long q_ts, long k_bs,
                     long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts, int B, int Hh,
                     int Lq, int Lk, int D, int DV, float scale, void* workspace, long workspace_bytes, void* stream);
int ms2_gemm_res_ln(const void* A, long lda, const void* W, const float* bias, const float* residual, long ldr,
                    float* x_out, long ldx, const float* gamma, const float* beta, float eps, void* t_out, int t_dt,
                    long ldt, int M, int K, void* stream);
int ms2_gemm_rope(const void* A, long lda, const void* W, const float* bias, void* out, long ldo, int M, int N, int K,
                  int L, int rope_cols, const float* cos_t, const float* sin_t, int table_len, void* stream);
int ms2_axpby(const float* x, float a, const float* z, float b, float c, void* y, int y_dt, long n, long zn, void* stream);
int ms2_bank_rows(const void* const* h_src, const void* const* h_pos, const long* h_pos_bs, const int* h_rows, int n, int W,
                  int B, void* k_in, long k_bs, void* m_out, long m_bs, int out_dt, void* stream);
}

// must mirror `struct ms2_memattn_layer_w` of include/medsam2_b200.h
struct LayerW {
  const float *norm1_g, *norm1_b, *norm2_g, *norm2_b, *norm3_g, *norm3_b;
  const void* qkv_w; const float* qkv_b;      // self-attention q|k|v fused [3C, C]
  const void* so_w; const float* so_b;        // self-attention out_proj [C, C]
  const void* cq_w; const float* cq_b;        // cross-attention q_proj [C, C]
  const void* ck_w; const float* ck_b;        // cross-attention k_proj [C, Cm]
  const void* vo_w; const float* vo_b;        // cross-attention out_proj . v_proj folded [C, Cm]
  const void* f1_w; const float* f1_b;        // FFN [F, C]
  const void* f2_w; const float* f2_b;        // FFN [C, F]
  const float *rope_cos, *rope_sin;           // axial RoPE table [L, D/2]
  float eps1, eps2, eps3;
  int C, Cm, F, rope_len;
};

namespace {

inline char* align_up(char* p) { return (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }

struct Ws {                 // carve-up of the caller's workspace for R = B*L rows
  void *t, *qkv, *o, *h, *q, *att;
  void* split; long split_bytes;
  bool ok;
};

Ws carve(void* ws, long ws_bytes, long R, int C, int Cm, int F) {
  Ws w;
  char* p = align_up((char*)ws);
  auto take = [&](long bytes) { void* r = p; p = align_up(p + bytes); return r; };
  w.t = take(R * C * 2);
  w.qkv = take(R * 3 * C * 2);
  w.o = take(R * C * 2);
  w.h = take(R * (long)F * 2);
  w.q = take(R * C * 2);
  w.att = take(R * Cm * 2);
  w.split = p;
  w.split_bytes = ws_bytes - (p - (char*)ws);
  w.ok = w.split_bytes >= 0;
  return w;
}

// the split-KV scratch the tensor-level wrappers (ops.attention / ops.attention_dv) would hand to the kernel: the same
// budget gives the same number of key splits, i.e. bit-identical results on both call paths
long split_budget(long avail, int B, int Lq, int Lk, int DV) {
  const long qtiles = (long)B * ((Lq + 127) / 128);
  if (qtiles >= 2 * 148 || Lk < 512) return 0;
  const long per_split = (long)B * Lq * (DV + 2) * 4;
  long nsplit = (4 * 148) / qtiles;
  if (nsplit > 32) nsplit = 32;
  if (nsplit > (256L << 20) / per_split) nsplit = (256L << 20) / per_split;
  if (nsplit < 2) nsplit = 2;
  const long want = nsplit * per_split;
  return want < avail ? want : avail;
}

#define MS2_TRY(call)       \
  do {                      \
    int rc__ = (call);      \
    if (rc__) return rc__;  \
  } while (0)

// Row-complete fused tiles (csrc/gemm_row.cu) for the shipped geometry: projection + RoPE and GEMM + residual + LayerNorm
// as one kernel each, 8 instead of 13 dependent kernels per layer.  OFF by default (MS2_MEMATTN_FUSED=1 enables): measured
// on the bench step 409 vs 421 slices/s — a 128 x 256 row-complete tile leaves 32 CTAs for 4096 rows, each of which has to
// pull its whole A and W panels through one SM's ~70 B/clk TMA ingest (1.5 MB for the FFN's second GEMM = 11 us), while
// the one-op-per-kernel chain spreads the same bytes over 128+ CTAs and overlaps its prologues through programmatic
// dependent launch (~4-5 us per kernel in the chain).
bool fused_ok(const LayerW& w, long R) {
  static const bool on = []() { const char* e = getenv("MS2_MEMATTN_FUSED"); return e && atoi(e) != 0; }();
  return on && w.C == 256 && R % 128 == 0 && w.Cm % 8 == 0 && w.Cm >= 16 && w.F % 8 == 0;
}

// x (fp32 [R,C], in place) through LN1 -> self-attention (RoPE on q,k) -> +residual -> LN2 -> q projection (+RoPE) -> q.
// t_ready: s.t already holds LN1(x) (written by the previous layer's fused FFN epilogue).
int layer_pre(const LayerW& w, float* x, void* q_out, const Ws& s, int B, int L, void* st, bool t_ready = false) {
  const int C = w.C;
  const long R = (long)B * L;
  const bool fused = fused_ok(w, R);
  if (!t_ready) MS2_TRY(ms2_layernorm(x, nullptr, w.norm1_g, w.norm1_b, s.t, MS2_BF16, (int)R, C, w.eps1, 0, st));
  bf16* qkv = (bf16*)s.qkv;
  if (fused) {
    MS2_TRY(ms2_gemm_rope(s.t, C, w.qkv_w, w.qkv_b, qkv, 3 * C, (int)R, 3 * C, C, L, 2 * C, w.rope_cos, w.rope_sin, w.rope_len, st));
  } else {
    MS2_TRY(ms2_gemm(s.t, MS2_BF16, C, w.qkv_w, MS2_BF16, w.qkv_b, nullptr, nullptr, 0, s.qkv, MS2_BF16, 3 * C, (int)R, 3 * C, C,
                     0, 0, st));
    if (B == 1) {   // q and k of the fused projection in one launch: "batch" 0 = q columns, 1 = k columns of the same rows
      MS2_TRY(ms2_rope(qkv, MS2_BF16, C, 3 * C, 2, L, L, C, w.rope_cos, w.rope_sin, w.rope_len, st));
    } else {
      MS2_TRY(ms2_rope(qkv, MS2_BF16, (long)L * 3 * C, 3 * C, B, L, L, C, w.rope_cos, w.rope_sin, w.rope_len, st));
      MS2_TRY(ms2_rope(qkv + C, MS2_BF16, (long)L * 3 * C, 3 * C, B, L, L, C, w.rope_cos, w.rope_sin, w.rope_len, st));
    }
  }
  MS2_TRY(ms2_attention_ws(qkv, qkv + C, qkv + 2 * C, s.o, MS2_BF16, (long)L * 3 * C, C, 3 * C, (long)L * 3 * C, C, 3 * C,
                           (long)L * 3 * C, C, 3 * C, (long)L * C, C, C, B, 1, L, L, C, 1.0f / sqrtf((float)C), 0, s.split,
                           split_budget(s.split_bytes, B, L, L, C), st));
  if (fused) {
    MS2_TRY(ms2_gemm_res_ln(s.o, C, w.so_w, w.so_b, x, C, x, C, w.norm2_g, w.norm2_b, w.eps2, s.t, MS2_BF16, C, (int)R, C, st));
    MS2_TRY(ms2_gemm_rope(s.t, C, w.cq_w, w.cq_b, q_out, C, (int)R, C, C, L, C, w.rope_cos, w.rope_sin, w.rope_len, st));
    return MS2_OK;
  }
  MS2_TRY(ms2_gemm(s.o, MS2_BF16, C, w.so_w, MS2_BF16, w.so_b, nullptr, x, C, x, MS2_F32, C, (int)R, C, C, 0, 0, st));
  MS2_TRY(ms2_layernorm(x, nullptr, w.norm2_g, w.norm2_b, s.t, MS2_BF16, (int)R, C, w.eps2, 0, st));
  MS2_TRY(ms2_gemm(s.t, MS2_BF16, C, w.cq_w, MS2_BF16, w.cq_b, nullptr, nullptr, 0, q_out, MS2_BF16, C, (int)R, C, C, 0, 0, st));
  MS2_TRY(ms2_rope(q_out, MS2_BF16, (long)L * C, C, B, L, L, C, w.rope_cos, w.rope_sin, w.rope_len, st));
  return MS2_OK;
}

// att (bf16 [R,Cm] = softmax(q K^T) M) -> x += att (Wo Wv)^T + b -> LN3 -> FFN(ReLU) -> +residual (x in place).
// next_g != null (fused geometry only): the LayerNorm that follows (the next layer's norm1, or the stack's final norm) is
// written by the FFN's epilogue into t_next (t_next_dt).
int layer_post(const LayerW& w, float* x, const void* att, const Ws& s, int B, int L, void* st, const float* next_g = nullptr,
               const float* next_b = nullptr, float next_eps = 0.f, void* t_next = nullptr, int t_next_dt = MS2_BF16) {
  const int C = w.C;
  const long R = (long)B * L;
  const bool fused = fused_ok(w, R);
  if (fused) {
    MS2_TRY(ms2_gemm_res_ln(att, w.Cm, w.vo_w, w.vo_b, x, C, x, C, w.norm3_g, w.norm3_b, w.eps3, s.t, MS2_BF16, C, (int)R, w.Cm, st));
  } else {
    MS2_TRY(ms2_gemm(att, MS2_BF16, w.Cm, w.vo_w, MS2_BF16, w.vo_b, nullptr, x, C, x, MS2_F32, C, (int)R, C, w.Cm, 0, 0, st));
    MS2_TRY(ms2_layernorm(x, nullptr, w.norm3_g, w.norm3_b, s.t, MS2_BF16, (int)R, C, w.eps3, 0, st));
  }
  MS2_TRY(ms2_gemm(s.t, MS2_BF16, C, w.f1_w, MS2_BF16, w.f1_b, nullptr, nullptr, 0, s.h, MS2_BF16, w.F, (int)R, w.F, C, 2, 0, st));
  if (fused && next_g)
    return ms2_gemm_res_ln(s.h, w.F, w.f2_w, w.f2_b, x, C, x, C, next_g, next_b, next_eps, t_next, t_next_dt, C, (int)R, w.F, st);
  MS2_TRY(ms2_gemm(s.h, MS2_BF16, w.F, w.f2_w, MS2_BF16, w.f2_b, nullptr, x, C, x, MS2_F32, C, (int)R, C, w.F, 0, 0, st));
  return MS2_OK;
}

}  // namespace

extern "C" long ms2_memattn_workspace_bytes(int B, int L, int C, int Cm, int F) {
  const long R = (long)B * L;
  const long fixed = R * C * 2 * 4 + R * 3 * C * 2 + R * (long)F * 2 + R * Cm * 2 + 8 * 256;
  // split-KV scratch of the self- and cross-attention kernels: up to 32 partials of [R, max(C, Cm) + 2] fp32
  const long per_split = R * (C + 2) * 4;
  long nsplit = 32;
  if (nsplit * per_split > (256L << 20)) nsplit = (256L << 20) / per_split;
  if (nsplit < 2) nsplit = 2;
  return fixed + nsplit * per_split;
}

extern "C" int ms2_memattn_layer_pre(const void* layer, float* x, void* q_out, void* ws, long ws_bytes, int B, int L,
                                     void* stream) {
  MS2_CHECK_ARG(layer && x && q_out && ws && B > 0 && L > 0, "memattn_layer_pre: bad args");
  const LayerW& w = *(const LayerW*)layer;
  const Ws s = carve(ws, ws_bytes, (long)B * L, w.C, w.Cm, w.F);
  MS2_CHECK_ARG(s.ok, "memattn_layer_pre: workspace too small (ms2_memattn_workspace_bytes)");
  return layer_pre(w, x, q_out, s, B, L, stream);
}

extern "C" int ms2_memattn_layer_post(const void* layer, float* x, const void* att, void* ws, long ws_bytes, int B, int L,
                                      void* stream) {
  MS2_CHECK_ARG(layer && x && att && ws && B > 0 && L > 0, "memattn_layer_post: bad args");
  const LayerW& w = *(const LayerW*)layer;
  const Ws s = carve(ws, ws_bytes, (long)B * L, w.C, w.Cm, w.F);
  MS2_CHECK_ARG(s.ok, "memattn_layer_post: workspace too small");
  return layer_post(w, x, att, s, B, L, stream);
}

// The per-frame rows of the memory bank (recent memories + object-pointer tokens): sources -> (mem + pos) and raw values
// (ms2_bank_rows), then per layer K = k_proj(mem + pos) written into the bank rows [row0, row0 + n) and rotated in place.
extern "C" int ms2_memattn_bank_project(const void* layers, int n_layers, const void* const* h_src, const void* const* h_pos,
                                        const long* h_pos_bs, const int* h_rows, int n_src, int B, int n_rope_rows, int Lq,
                                        void* const* h_K, long k_bs, int row0, void* m_bank, long m_bs, void* ws,
                                        long ws_bytes, void* stream) {
  MS2_CHECK_ARG(layers && n_layers > 0 && h_src && h_rows && h_K && m_bank && ws && B > 0, "memattn_bank_project: bad args");
  const LayerW* w = (const LayerW*)layers;
  const int C = w[0].C, Cm = w[0].Cm;
  long n = 0;
  for (int i = 0; i < n_src; ++i) n += h_rows[i];
  if (!n) return MS2_OK;
  MS2_CHECK_ARG(ws_bytes >= (long)B * n * Cm * 2 + 256, "memattn_bank_project: workspace too small");
  void* k_in = align_up((char*)ws);
  for (int c0 = 0, r0 = 0; c0 < n_src; c0 += 80) {          // ms2_bank_rows takes <= 80 sources per launch
    const int cn = n_src - c0 < 80 ? n_src - c0 : 80;
    long rows_c = 0;
    for (int i = 0; i < cn; ++i) rows_c += h_rows[c0 + i];
    MS2_TRY(ms2_bank_rows(h_src + c0, h_pos ? h_pos + c0 : nullptr, h_pos_bs ? h_pos_bs + c0 : nullptr, h_rows + c0, cn, Cm, B,
                          (bf16*)k_in + (long)r0 * Cm, n * Cm, (bf16*)m_bank + ((long)row0 + r0) * Cm, m_bs, MS2_BF16, stream));
    r0 += (int)rows_c;
  }
  for (int l = 0; l < n_layers; ++l) {
    for (int b = 0; b < B; ++b) {
      bf16* kd = (bf16*)h_K[l] + (long)b * k_bs + (long)row0 * C;
      MS2_TRY(ms2_gemm((const bf16*)k_in + (long)b * n * Cm, MS2_BF16, Cm, w[l].ck_w, MS2_BF16, w[l].ck_b, nullptr, nullptr, 0,
                       kd, MS2_BF16, C, (int)n, C, Cm, 0, 0, stream));
      if (n_rope_rows > 0)
        MS2_TRY(ms2_rope(kd, MS2_BF16, n * C, C, 1, (int)n, n_rope_rows, C, w[l].rope_cos, w[l].rope_sin, w[l].rope_len, stream));
    }
  }
  (void)Lq;
  return MS2_OK;
}

// The whole stack on one GPU: x = curr + 0.1 * curr_pos; per layer pre -> cross-attention over the bank (K_l, raw values M)
// -> post; final LayerNorm -> out (fp32 [B,L,C]).
extern "C" int ms2_memattn_forward(const void* layers, int n_layers, const float* curr, const float* curr_pos, float pos_scale,
                                   void* const* h_K, long k_bs, const void* m_bank, long m_bs, int Lk, const float* norm_g,
                                   const float* norm_b, float norm_eps, float* x, float* out, void* ws, long ws_bytes, int B,
                                   int L, void* stream) {
  MS2_CHECK_ARG(layers && n_layers > 0 && curr && h_K && m_bank && x && out && ws && B > 0 && L > 0 && Lk > 0,
                "memattn_forward: bad args");
  const LayerW* w = (const LayerW*)layers;
  const int C = w[0].C, Cm = w[0].Cm;
  const long R = (long)B * L;
  const Ws s = carve(ws, ws_bytes, R, C, Cm, w[0].F);
  MS2_CHECK_ARG(s.ok, "memattn_forward: workspace too small (ms2_memattn_workspace_bytes)");
  if (curr_pos) MS2_TRY(ms2_axpby(curr, 1.0f, curr_pos, pos_scale, 0.0f, x, MS2_F32, R * C, R * C, stream));
  else MS2_CUDA(cudaMemcpyAsync(x, curr, R * C * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "memattn_forward copy");
  const bool fused = fused_ok(w[0], R) && out != x;
  for (int l = 0; l < n_layers; ++l) {
    MS2_TRY(layer_pre(w[l], x, s.q, s, B, L, stream, fused && l > 0));
    MS2_TRY(ms2_attention_dv(s.q, h_K[l], m_bank, s.att, MS2_BF16, (long)L * C, C, C, k_bs, C, C, m_bs, Cm, Cm, (long)L * Cm, Cm,
                             Cm, B, 1, L, Lk, C, Cm, 1.0f / sqrtf((float)C), s.split, split_budget(s.split_bytes, B, L, Lk, Cm),
                             stream));
    if (!fused) MS2_TRY(layer_post(w[l], x, s.att, s, B, L, stream));
    else if (l + 1 < n_layers)      // the FFN epilogue writes the next layer's LN1 (bf16) ...
      MS2_TRY(layer_post(w[l], x, s.att, s, B, L, stream, w[l + 1].norm1_g, w[l + 1].norm1_b, w[l + 1].eps1, s.t, MS2_BF16));
    else                            // ... or the stack's final LayerNorm (fp32) straight into `out`
      return layer_post(w[l], x, s.att, s, B, L, stream, norm_g, norm_b, norm_eps, out, MS2_F32);
  }
  return ms2_layernorm(x, nullptr, norm_g, norm_b, out, MS2_F32, (int)R, C, norm_eps, 0, stream);
}
