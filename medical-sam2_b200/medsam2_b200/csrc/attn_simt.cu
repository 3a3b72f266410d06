// Reference-precision flash attention on the CUDA cores (fp32 accumulate, online softmax):
//   * dense mode      — F.scaled_dot_product_attention call sites of transformer.py:252-258,318
//   * windowed mode   — Hiera MultiScaleAttention (hieradet.py:58-83) with window partition,
//                       zero-pad-AFTER-norm semantics (pad tokens are keys with k=b_k, v=b_v; SURVEY
//                       §0 finding 6 / App. A.3), 2x2 q max-pool and unpartition+crop folded into
//                       the gather/scatter index math, so no padded copy of the tokens ever exists.
// This is the exact ("fp32 mode") path and the on-device cross-check of the tcgen05 kernels.
#include "common.cuh"

namespace {

struct AttnP {
  const void *q, *k, *v;
  void* o;
  long q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts;
  int B, Hh, Lq, Lk;
  float scale;
  int win;  // 0 dense, 1 windowed
  int H, W, ws, nwx, nwy, qpool, Ho, Wo, dim_out;
  const float* bias;
};

template <typename T, int D>
struct Loader {
  const AttnP& p;
  int bz, h, b, wy, wx;
  __device__ Loader(const AttnP& p_, int bz_, int h_) : p(p_), bz(bz_), h(h_) {
    if (p.win) {
      int nw = p.nwx * p.nwy;
      b = bz / nw;
      int w = bz - b * nw;
      wy = w / p.nwx;
      wx = w - wy * p.nwx;
    } else {
      b = bz; wy = wx = 0;
    }
  }
  // value of q/k/v (which = 0/1/2) at padded-grid position (y,x); pad -> bias
  __device__ __forceinline__ float win_val(int which, int y, int x, int d) const {
    const int col = which * p.dim_out + h * D + d;
    if (y < p.H && x < p.W)
      return to_f(((const T*)p.q)[(((long)b * p.H + y) * p.W + x) * (3L * p.dim_out) + col]);
    return p.bias[col];
  }
  __device__ __forceinline__ float k_at(int j, int d) const {
    if (!p.win) return to_f(((const T*)p.k)[bz * p.k_bs + h * p.k_hs + (long)j * p.k_ts + d]);
    return win_val(1, wy * p.ws + j / p.ws, wx * p.ws + j % p.ws, d);
  }
  __device__ __forceinline__ float v_at(int j, int d) const {
    if (!p.win) return to_f(((const T*)p.v)[bz * p.v_bs + h * p.v_hs + (long)j * p.v_ts + d]);
    return win_val(2, wy * p.ws + j / p.ws, wx * p.ws + j % p.ws, d);
  }
  __device__ __forceinline__ float q_at(int i, int d) const {
    if (!p.win) return to_f(((const T*)p.q)[bz * p.q_bs + h * p.q_hs + (long)i * p.q_ts + d]);
    if (!p.qpool) return win_val(0, wy * p.ws + i / p.ws, wx * p.ws + i % p.ws, d);
    const int hw = p.ws >> 1;
    const int y = wy * p.ws + 2 * (i / hw), x = wx * p.ws + 2 * (i % hw);
    return fmaxf(fmaxf(win_val(0, y, x, d), win_val(0, y, x + 1, d)),
                 fmaxf(win_val(0, y + 1, x, d), win_val(0, y + 1, x + 1, d)));
  }
  // output element offset for query i, or -1 when the query falls in the cropped padding
  __device__ __forceinline__ long o_off(int i) const {
    if (!p.win) return bz * p.o_bs + h * p.o_hs + (long)i * p.o_ts;
    const int w = p.qpool ? (p.ws >> 1) : p.ws;
    const int y = wy * w + i / w, x = wx * w + i % w;
    if (y >= p.Ho || x >= p.Wo) return -1;
    return (((long)b * p.Ho + y) * p.Wo + x) * (long)p.dim_out + h * D;
  }
};

template <typename T, int D, int BK, int R>
__global__ void __launch_bounds__(128)
attn_simt_kernel(const AttnP p) {
  MS2_PDL_WAIT();
  constexpr int KPL = BK / 32;
  constexpr int DI = (D + 31) / 32;
  constexpr int NW = 4;
  extern __shared__ float sm[];
  float* Ks = sm;                   // [BK][D+1]
  float* Vs = Ks + BK * (D + 1);    // [BK][D]
  float* Qs = Vs + BK * D;          // [NW][R][D]
  float* Ps = Qs + NW * R * D;      // [NW][R][BK]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Loader<T, D> ld(p, blockIdx.z, blockIdx.y);
  const int q0 = blockIdx.x * (NW * R) + warp * R;
  long ooff[R];
  float m[R], l[R], acc[R][DI];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int qi = q0 + r;
    ooff[r] = (qi < p.Lq) ? ld.o_off(qi) : -1;
    for (int d = lane; d < D; d += 32) Qs[(warp * R + r) * D + d] = (ooff[r] >= 0) ? ld.q_at(qi, d) * p.scale : 0.f;
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DI; ++i) acc[r][i] = 0.f;
  }
  for (int kt = 0; kt < p.Lk; kt += BK) {
    __syncthreads();
    for (int idx = tid; idx < BK * D; idx += 128) {
      const int j = idx / D, d = idx - j * D;
      const int kj = kt + j;
      float kv = 0.f, vv = 0.f;
      if (kj < p.Lk) { kv = ld.k_at(kj, d); vv = ld.v_at(kj, d); }
      Ks[j * (D + 1) + d] = kv;
      Vs[j * D + d] = vv;
    }
    __syncthreads();
    float s[R][KPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < KPL; ++c) s[r][c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float kv[KPL];
#pragma unroll
      for (int c = 0; c < KPL; ++c) kv[c] = Ks[(lane + 32 * c) * (D + 1) + d];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float qv = Qs[(warp * R + r) * D + d];
#pragma unroll
        for (int c = 0; c < KPL; ++c) s[r][c] = fmaf(qv, kv[c], s[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < KPL; ++c) {
        if (kt + lane + 32 * c >= p.Lk) s[r][c] = -INFINITY;
        mx = fmaxf(mx, s[r][c]);
      }
      mx = warp_max(mx);
      const float m_new = fmaxf(m[r], mx);
      const float corr = expf(m[r] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < KPL; ++c) {
        const float pv = expf(s[r][c] - m_new);
        ps += pv;
        Ps[(warp * R + r) * BK + lane + 32 * c] = pv;
      }
      ps = warp_sum(ps);
      l[r] = l[r] * corr + ps;
      m[r] = m_new;
#pragma unroll
      for (int i = 0; i < DI; ++i) acc[r][i] *= corr;
    }
    __syncwarp();
    for (int j = 0; j < BK; ++j) {
      float vv[DI];
#pragma unroll
      for (int i = 0; i < DI; ++i) vv[i] = (lane + 32 * i < D) ? Vs[j * D + lane + 32 * i] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float pj = Ps[(warp * R + r) * BK + j];
#pragma unroll
        for (int i = 0; i < DI; ++i) acc[r][i] = fmaf(pj, vv[i], acc[r][i]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (ooff[r] < 0) continue;
    const float inv = 1.f / l[r];
    T* o = (T*)p.o + ooff[r];
#pragma unroll
    for (int i = 0; i < DI; ++i)
      if (lane + 32 * i < D) o[lane + 32 * i] = from_f<T>(acc[r][i] * inv);
  }
}

template <typename T, int D, int BK>
int launch_d(const AttnP& p, int nbatch, cudaStream_t st) {
  constexpr int R = 4;
  size_t smem = sizeof(float) * ((size_t)BK * (D + 1) + (size_t)BK * D + 4 * R * D + 4 * R * BK);
  auto kern = attn_simt_kernel<T, D, BK, R>;
  MS2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attn attr");
  dim3 grid(ceil_div(p.Lq, 4 * R), p.Hh, nbatch);
  ms2_launch(kern, grid, 128, smem, st, p);
  MS2_CHECK_LAUNCH("attn_simt_kernel");
  return MS2_OK;
}

template <typename T>
int launch_t(const AttnP& p, int D, int nbatch, cudaStream_t st) {
  switch (D) {
    case 16: return launch_d<T, 16, 64>(p, nbatch, st);
    case 32: return launch_d<T, 32, 64>(p, nbatch, st);
    case 64: return launch_d<T, 64, 64>(p, nbatch, st);
    case 96: return launch_d<T, 96, 64>(p, nbatch, st);
    case 128: return launch_d<T, 128, 32>(p, nbatch, st);
    case 256: return launch_d<T, 256, 32>(p, nbatch, st);
  }
  ms2_set_error("attention: unsupported head dim %d", D);
  return MS2_ERR_UNSUPPORTED;
}

}  // namespace

int ms2_attention_tc_launch(const void* q, const void* k, const void* v, void* o, long q_bs, long q_hs, long q_ts,
                            long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs,
                            long o_ts, int B, int Hh, int Lq, int Lk, int D, int DV, float scale, void* ws,
                            long ws_bytes, cudaStream_t st);
bool ms2_attention_tc_supported(int dt, long q_hs, long q_ts, long k_hs, long k_ts, long v_hs, long v_ts, long o_hs,
                                long o_ts, int Hh, int Lq, int Lk, int D, int DV);

int ms2_attention_small(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs, long q_ts,
                        long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs,
                        long o_ts, int B, int Hh, int Lq, int Lk, int D, float scale, void* ws, long ws_bytes,
                        cudaStream_t st);

extern "C" int ms2_attention_dv(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs,
                                long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs,
                                long o_hs, long o_ts, int B, int Hh, int Lq, int Lk, int D, int DV, float scale,
                                void* workspace, long workspace_bytes, void* stream) {
  MS2_CHECK_ARG(q && k && v && o, "attention: null pointer");
  MS2_CHECK_ARG(B >= 0 && Hh > 0 && Lq >= 0 && Lk > 0, "attention: bad shape");
  if (B == 0 || Lq == 0) return MS2_OK;
  MS2_CHECK_ARG(ms2_attention_tc_supported(dt, q_hs, q_ts, k_hs, k_ts, v_hs, v_ts, o_hs, o_ts, Hh, Lq, Lk, D, DV),
                "attention_dv: unsupported shape/dtype/stride (bf16, D=256, DV=64, Lq,Lk >= 64)");
  return ms2_attention_tc_launch(q, k, v, o, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts, B, Hh,
                                 Lq, Lk, D, DV, scale, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ms2_attention_dv_partial_launch(const void* q, const void* k, const void* v, float* part_o, float* part_ml, long q_bs,
                                    long q_ts, long k_bs, long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk,
                                    float scale, void* ws, long ws_bytes, cudaStream_t st);
int ms2_attention_merge_launch(const float* parts_o, const float* parts_ml, long part_stride, void* o, long o_bs, long o_ts,
                               int B, int Lq, int nparts, cudaStream_t st);

extern "C" int ms2_attention_dv_partial(const void* q, const void* k, const void* v, float* part_o, float* part_ml, int dt,
                                        long q_bs, long q_ts, long k_bs, long k_ts, long v_bs, long v_ts, int B, int Lq,
                                        int Lk, int D, int DV, float scale, void* workspace, long workspace_bytes,
                                        void* stream) {
  MS2_CHECK_ARG(q && k && v && part_o && part_ml, "attention_dv_partial: null pointer");
  MS2_CHECK_ARG(B > 0 && Lq > 0 && Lk > 0, "attention_dv_partial: bad shape");
  MS2_CHECK_ARG(ms2_attention_tc_supported(dt, 256, q_ts, 256, k_ts, 64, v_ts, 64, 64, 1, Lq, Lk, D, DV) && D == 256 &&
                    DV == 64,
                "attention_dv_partial: unsupported shape/dtype/stride (bf16, D=256, DV=64, Lq,Lk >= 64)");
  return ms2_attention_dv_partial_launch(q, k, v, part_o, part_ml, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, B, Lq, Lk, scale,
                                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int ms2_attention_merge(const float* parts_o, const float* parts_ml, long part_stride, void* o, int dt,
                                   long o_bs, long o_ts, int B, int Lq, int DV, int nparts, void* stream) {
  MS2_CHECK_ARG(parts_o && parts_ml && o && nparts >= 1, "attention_merge: bad args");
  MS2_CHECK_ARG(dt == MS2_BF16 && DV == 64, "attention_merge: bf16 output with DV = 64 only");
  if (B == 0 || Lq == 0) return MS2_OK;
  return ms2_attention_merge_launch(parts_o, parts_ml, part_stride, o, o_bs, o_ts, B, Lq, nparts, (cudaStream_t)stream);
}

int ms2_attention_dv_partial_push_launch(const void* q, const void* k, const void* v, long q_bs, long q_ts, long k_bs,
                                         long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk, float scale, void* ws,
                                         long ws_bytes, const void* const* h_dst, const void* const* h_flag, int world,
                                         unsigned step, void* counter, cudaStream_t st);
int ms2_attention_merge_wait_launch(const float* parts, long part_stride, const void* flags, unsigned step, int world, void* o,
                                    long o_bs, long o_ts, int B, int Lq, cudaStream_t st);

extern "C" int ms2_attention_dv_partial_push(const void* q, const void* k, const void* v, int dt, long q_bs, long q_ts,
                                             long k_bs, long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk, int D,
                                             int DV, float scale, void* workspace, long workspace_bytes,
                                             const void* const* h_dst, const void* const* h_flag, int world, int step,
                                             void* counter, void* stream) {
  MS2_CHECK_ARG(B > 0 && Lq > 0 && Lk >= 0 && D == 256 && DV == 64 && dt == MS2_BF16,
                "attention_dv_partial_push: bf16, D=256, DV=64 only");
  MS2_CHECK_ARG(Lk == 0 || (q && k && v), "attention_dv_partial_push: null pointer");
  MS2_CHECK_ARG(Lk == 0 || ms2_attention_tc_supported(dt, 256, q_ts, 256, k_ts, 64, v_ts, 64, 64, 1, Lq, Lk, D, DV),
                "attention_dv_partial_push: unsupported shape/stride (Lq, Lk >= 64)");
  return ms2_attention_dv_partial_push_launch(q, k, v, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, B, Lq, Lk, scale, workspace,
                                              workspace_bytes, h_dst, h_flag, world, (unsigned)step, counter,
                                              (cudaStream_t)stream);
}

extern "C" int ms2_attention_merge_wait(const float* parts, long part_stride, const void* flags, int step, int world, void* o,
                                        int dt, long o_bs, long o_ts, int B, int Lq, int DV, void* stream) {
  MS2_CHECK_ARG(dt == MS2_BF16 && DV == 64, "attention_merge_wait: bf16 output with DV = 64 only");
  if (B == 0 || Lq == 0) return MS2_OK;
  return ms2_attention_merge_wait_launch(parts, part_stride, flags, (unsigned)step, world, o, o_bs, o_ts, B, Lq,
                                         (cudaStream_t)stream);
}

extern "C" int ms2_attention_ws(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs,
                                long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs,
                                long o_hs, long o_ts, int B, int Hh, int Lq, int Lk, int D, float scale, int impl,
                                void* workspace, long workspace_bytes, void* stream) {
  MS2_CHECK_ARG(q && k && v && o, "attention: null pointer");
  MS2_CHECK_ARG(B >= 0 && Hh > 0 && Lq >= 0 && Lk > 0, "attention: bad shape");
  if (B == 0 || Lq == 0) return MS2_OK;
  const bool tc_ok = ms2_attention_tc_supported(dt, q_hs, q_ts, k_hs, k_ts, v_hs, v_ts, o_hs, o_ts, Hh, Lq, Lk, D, D);
  if (impl == 2) MS2_CHECK_ARG(tc_ok, "attention: tcgen05 path does not support this shape/dtype/stride");
  if (impl == 0 || impl == 3) {
    const int rc = ms2_attention_small(q, k, v, o, dt, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs,
                                       o_ts, B, Hh, Lq, Lk, D, scale, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc < 0) return rc;
    if (rc == 1) return MS2_OK;
    MS2_CHECK_ARG(impl != 3, "attention: small-shape path does not support this configuration");
  }
  if (impl == 2 || (impl == 0 && tc_ok))
    return ms2_attention_tc_launch(q, k, v, o, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts, B,
                                   Hh, Lq, Lk, D, D, scale, workspace, workspace_bytes, (cudaStream_t)stream);
  AttnP p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.k = k; p.v = v; p.o = o;
  p.q_bs = q_bs; p.q_hs = q_hs; p.q_ts = q_ts;
  p.k_bs = k_bs; p.k_hs = k_hs; p.k_ts = k_ts;
  p.v_bs = v_bs; p.v_hs = v_hs; p.v_ts = v_ts;
  p.o_bs = o_bs; p.o_hs = o_hs; p.o_ts = o_ts;
  p.B = B; p.Hh = Hh; p.Lq = Lq; p.Lk = Lk; p.scale = scale; p.win = 0;
  MS2_DISPATCH_DTYPE(dt, T, return launch_t<T>(p, D, B, (cudaStream_t)stream));
  return MS2_OK;
}

extern "C" int ms2_attention(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs,
                             long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs,
                             long o_hs, long o_ts, int B, int Hh, int Lq, int Lk, int D, float scale, int impl,
                             void* stream) {
  return ms2_attention_ws(q, k, v, o, dt, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts, B, Hh,
                          Lq, Lk, D, scale, impl, nullptr, 0, stream);
}

bool ms2_window_attention_tc_supported(int dt, int heads, int D, int ws, int qpool);
int ms2_window_attention_tc_launch(const void* qkv, const float* qkv_bias, void* out, int B, int H, int W, int heads,
                                   int ws, int qpool, float scale, cudaStream_t st);

extern "C" int ms2_window_attention_impl(const void* qkv, const float* qkv_bias, void* out, int dt, int B, int H, int W,
                                         int heads, int D, int ws, int qpool, float scale, int impl, void* stream) {
  MS2_CHECK_ARG(qkv && qkv_bias && out, "window_attention: null pointer");
  MS2_CHECK_ARG(ws > 0 && (!qpool || ws % 2 == 0), "window_attention: bad window %d", ws);
  MS2_CHECK_ARG(!qpool || (H % 2 == 0 && W % 2 == 0), "window_attention: q-pool needs even H,W");
  if (B == 0) return MS2_OK;
  const bool tc_ok = ms2_window_attention_tc_supported(dt, heads, D, ws, qpool);
  if (impl == 2) MS2_CHECK_ARG(tc_ok, "window_attention: tcgen05 path does not support this configuration");
  if (impl == 2 || (impl == 0 && tc_ok))
    return ms2_window_attention_tc_launch(qkv, qkv_bias, out, B, H, W, heads, ws, qpool, scale, (cudaStream_t)stream);
  AttnP p;
  memset(&p, 0, sizeof(p));
  p.q = qkv; p.k = qkv; p.v = qkv; p.o = out; p.bias = qkv_bias;
  p.B = B; p.Hh = heads; p.scale = scale; p.win = 1;
  p.H = H; p.W = W; p.ws = ws; p.qpool = qpool;
  p.nwy = (H + ws - 1) / ws; p.nwx = (W + ws - 1) / ws;
  p.Ho = qpool ? H / 2 : H; p.Wo = qpool ? W / 2 : W;
  p.dim_out = heads * D;
  p.Lk = ws * ws;
  p.Lq = qpool ? (ws / 2) * (ws / 2) : ws * ws;
  const int nbatch = B * p.nwy * p.nwx;
  MS2_DISPATCH_DTYPE(dt, T, return launch_t<T>(p, D, nbatch, (cudaStream_t)stream));
  return MS2_OK;
}

extern "C" int ms2_window_attention(const void* qkv, const float* qkv_bias, void* out, int dt, int B, int H, int W,
                                    int heads, int D, int ws, int qpool, float scale, void* stream) {
  return ms2_window_attention_impl(qkv, qkv_bias, out, dt, B, H, W, heads, D, ws, qpool, scale, 0, stream);
}
