// Latency-oriented kernels for the token side of the SAM mask decoder ("many tiny
// kernels on GPU, launch-bound"): these shapes have a handful of rows and cannot fill a tensor-core tile,
// so they are HBM/L2-latency bound and are written as coalesced, vectorised SIMT kernels.
//   * gemm_smallm : out[M<=16 per pass, N] = epilogue(A @ W^T); 8 lanes share one output column and split K
//                   in 16-byte chunks, A rows come from L1 (a few KB), W streams once.
//   * attn_fewk   : Lk <= 32 keys (image -> token attention, transformer.py:183-189): one thread per
//                   (query, head), K/V of all heads staged in shared memory as fp32.
//   * attn_fewq   : Lq <= 16 queries over any Lk that fits shared memory (token -> image attention,
//                   transformer.py:167-181,110-116): one CTA per (head, batch); scores for all keys are
//                   kept in shared memory (exact two-pass softmax), P V is reduced across 16 warps.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ small-M GEMM
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&f)[8]) {
    const uint4 u = *(const uint4*)p;
    const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

constexpr int SM_MT = 16;       // rows per pass
constexpr int SM_WARPS = 4;     // grouped kernel: 4 warps x 4 columns = 16 columns per CTA

// Stage rows [m0, m0+mt) of A (K elements each) in shared memory with asynchronous 16-byte copies: the row loop used
// to fetch its activations row by row from global memory, each row behind a branch - one L2 round trip PER ROW
// (measured 3.3 us for M = 1 vs 7.8 us for M = 9 at N = K = 256).
template <typename T>
__device__ __forceinline__ void smallm_stage_a(const T* __restrict__ A, long lda, int K, int m0, int mt, T* sA) {
  constexpr int VE = 16 / sizeof(T);                         // elements per 16-byte vector
  const int vpr = K / VE;
  __syncthreads();                                           // previous pass has finished reading sA
  for (int r = 0; r < mt; ++r) {
    const T* src = A + (long)(m0 + r) * lda;
    T* dst = sA + (long)r * K;
    for (int c = threadIdx.x; c < vpr; c += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + c * VE)),
                   "l"(src + c * VE) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void smallm_stage_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

// acc[m] += sum_k sA[m, k] * wrow[k] over this lane's 16-byte chunks (chunk stride KL*8 elements), four chunks per
// pass.  Branch-free: rows beyond mt re-read the last row and chunks beyond K multiply by a zero weight, so every
// shared-memory load of a pass is independent and the compiler batches them.
// All weight chunks of this lane (chunk c covers k = (kl + c*KL)*8 .. +8; NCH <= 8 chunks, zero beyond K) are loaded
// ONCE, before the wait for the staged activations, so the weight round trip (HBM, once per slice) overlaps the
// activation round trip.
template <typename T, int KL, int NCH>
__device__ __forceinline__ void smallm_load_w(const T* __restrict__ wrow, int K, int kl, float (&w)[NCH][8], int (&kc)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int k = (kl + c * KL) * 8;
    kc[c] = k < K ? k : K - 8;
    Vec8<T>::load(wrow + kc[c], w[c]);
    if (k >= K) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[c][i] = 0.f;
    }
  }
}

// One output row per iteration of a ROLLED loop: the fully unrolled 16-row version was ~2400 instructions (38 KB) that
// every warp fetched exactly once - ncu showed the kernel waiting for instructions (stall_no_inst), not for data.
template <typename T, typename TO, int KL, int NCH>
__device__ __forceinline__ void smallm_rows(const T* sA, int K, int kl, bool n_ok, int n, int m0, int mt,
                                            const float (&w)[NCH][8], const int (&kc)[NCH], float bv, float cs,
                                            const float* __restrict__ residual, long ldr, TO* __restrict__ out, long ldo,
                                            int act) {
  constexpr int R = 3;                                   // rows in flight per iteration (independent chains)
#pragma unroll 1
  for (int m = 0; m < mt; m += R) {
    float v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const T* arow = sA + (long)(m + r < mt ? m + r : mt - 1) * K;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float a[8];
        Vec8<T>::load(arow + kc[c], a);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          acc0 = fmaf(a[i], w[c][i], acc0);
          acc1 = fmaf(a[i + 1], w[c][i + 1], acc1);
        }
      }
      v[r] = acc0 + acc1;
    }
#pragma unroll
    for (int o = KL / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] += __shfl_xor_sync(0xffffffffu, v[r], o);
    // lane kl finishes row m + kl (kl < R): the row stores of a column go out in parallel
    float mine = v[0];
#pragma unroll
    for (int r = 1; r < R; ++r) mine = kl == r ? v[r] : mine;
    if (kl < R && m + kl < mt && n_ok) {
      float x = mine + bv;
      if (act == 1) x = gelu_erf(x);
      else if (act == 2) x = fmaxf(x, 0.f);
      else if (act == 3) x = 1.f / (1.f + __expf(-x));
      x *= cs;
      if (residual) x += residual[(long)(m0 + m + kl) * ldr + n];
      out[(long)(m0 + m + kl) * ldo + n] = from_f<TO>(x);
    }
  }
}

// KL lanes share one output column and split K in 16-byte chunks (KL = 8: 4 columns per warp; KL = 32: one column
// per warp, for K >= 1024 where 8 lanes would each walk 128+ elements serially).  Two warps per CTA so that even
// N = 256 gives 32..128 CTAs: these GEMMs are pure latency (W is 0.1-1 MB, A a few KB).
constexpr int SMK_WARPS = 2;
template <typename T, typename TO, int KL, int NCH>
__global__ void __launch_bounds__(SMK_WARPS * 32)
gemm_smallm_kernel(const T* __restrict__ A, long lda, const T* __restrict__ W, const float* __restrict__ bias,
                   const float* __restrict__ colscale, const float* __restrict__ residual, long ldr,
                   TO* __restrict__ out, long ldo, int M, int N, int K, int act, int rows_per_pass) {
  MS2_PDL_WAIT();
  constexpr int CPW = 32 / KL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = lane / KL, kl = lane % KL;
  const int n = (blockIdx.x * SMK_WARPS + warp) * CPW + cg;
  const bool n_ok = n < N;
  const T* wrow = W + (long)(n_ok ? n : 0) * K;
  extern __shared__ uint4 smallm_smem[];
  T* sA = (T*)smallm_smem;
  for (int m0 = 0; m0 < M; m0 += rows_per_pass) {
    const int mt = min(rows_per_pass, M - m0);
    smallm_stage_a<T>(A, lda, K, m0, mt, sA);
    float w[NCH][8];
    int kc[NCH];
    smallm_load_w<T, KL, NCH>(wrow, K, kl, w, kc);
    const float bv = (bias && n_ok) ? bias[n] : 0.f, cs = (colscale && n_ok) ? colscale[n] : 1.f;
    smallm_stage_wait();
    smallm_rows<T, TO, KL, NCH>(sA, K, kl, n_ok, n, m0, mt, w, kc, bv, cs, residual, ldr, out, ldo, act);
  }
}

// ---- grouped variant: up to 8 independent small-M GEMMs of the same M and K in ONE launch (blockIdx.y = group):
//      the 6 three-layer head MLPs of the mask decoder (4 hyper-networks, IoU, object score; mask_decoder.py:75-92,
//      238-267) run layer by layer as 3 launches instead of 18.
constexpr int SM_MAXG = 8;
template <typename T>
struct GroupedP {
  const T* A[SM_MAXG];
  const T* W[SM_MAXG];
  const float* bias[SM_MAXG];
  void* out[SM_MAXG];
  long lda[SM_MAXG], ldo[SM_MAXG];
  int N[SM_MAXG], act[SM_MAXG];
};

template <typename T, typename TO, int NCH>
__global__ void __launch_bounds__(SM_WARPS * 32)
gemm_smallm_grouped_kernel(const GroupedP<T> p, int M, int K, int rows_per_pass) {
  MS2_PDL_WAIT();
  const int g = blockIdx.y;
  const int N = p.N[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = lane >> 3, kl = lane & 7;
  const int n = (blockIdx.x * SM_WARPS + warp) * 4 + cg;
  if ((int)(blockIdx.x * SM_WARPS * 4) >= N) return;
  const bool n_ok = n < N;
  const T* A = p.A[g];
  const long lda = p.lda[g];
  const T* wrow = p.W[g] + (long)(n_ok ? n : 0) * K;
  const float* bias = p.bias[g];
  TO* out = (TO*)p.out[g];
  const long ldo = p.ldo[g];
  const int act = p.act[g];
  extern __shared__ uint4 smallm_smem[];
  T* sA = (T*)smallm_smem;
  for (int m0 = 0; m0 < M; m0 += rows_per_pass) {
    const int mt = min(rows_per_pass, M - m0);
    smallm_stage_a<T>(A, lda, K, m0, mt, sA);
    float w[NCH][8];
    int kc[NCH];
    smallm_load_w<T, 8, NCH>(wrow, K, kl, w, kc);
    const float bv = (bias && n_ok) ? bias[n] : 0.f;
    smallm_stage_wait();
    smallm_rows<T, TO, 8, NCH>(sA, K, kl, n_ok, n, m0, mt, w, kc, bv, 1.f, nullptr, 0, out, ldo, act);
  }
}

// ------------------------------------------------------------------ attention, few keys
template <typename T, int D>
__global__ void __launch_bounds__(256)
attn_fewk_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                 long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts,
                 long o_bs, long o_hs, long o_ts, int Hh, int Lq, int Lk, float scale) {
  MS2_PDL_WAIT();
  extern __shared__ float smf[];
  const int HS = Lk * D + 4;             // head stride, padded: heads land in different banks
  float* Ks = smf;                       // [Hh][Lk][D] (+4 per head)
  float* Vs = smf + Hh * HS;
  const int b = blockIdx.y;
  // 16-byte vectors, two per thread in flight (the scalar loop paid one global round trip per element and pass)
  constexpr int VPR = D / 8;
  const int nvec = Hh * Lk * VPR;
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 2 * blockDim.x) {
    float kk[2][8], vv[2][8];
    int dst[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      dst[u] = -1;
      if (i < nvec) {
        const int d8 = i % VPR, j = (i / VPR) % Lk, h = i / (VPR * Lk);
        Vec8<T>::load(k + b * k_bs + h * k_hs + (long)j * k_ts + d8 * 8, kk[u]);
        Vec8<T>::load(v + b * v_bs + h * v_hs + (long)j * v_ts + d8 * 8, vv[u]);
        dst[u] = h * HS + j * D + d8 * 8;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (dst[u] >= 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          Ks[dst[u] + e] = kk[u][e];
          Vs[dst[u] + e] = vv[u][e];
        }
      }
  }
  __syncthreads();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= Lq * Hh) return;
  const int h = gid % Hh, qi = gid / Hh;
  float qv[D];
  const T* qp = q + b * q_bs + h * q_hs + (long)qi * q_ts;
#pragma unroll
  for (int d = 0; d < D; d += 8) {
    float t[8];
    Vec8<T>::load(qp + d, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[d + i] = t[i] * scale;
  }
  const float* kh = Ks + h * HS;
  const float* vh = Vs + h * HS;
  float mx = -INFINITY, l = 0.f, acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
#pragma unroll 1
  for (int j = 0; j < Lk; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) a = fmaf(qv[d], kh[j * D + d], a);
    const float mn = fmaxf(mx, a);
    const float corr = __expf(mx - mn), pj = __expf(a - mn);
    mx = mn;
    l = l * corr + pj;
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = fmaf(pj, vh[j * D + d], acc[d] * corr);
  }
  const float inv = 1.f / l;
  T* op = o + b * o_bs + h * o_hs + (long)qi * o_ts;
#pragma unroll
  for (int d = 0; d < D; ++d) op[d] = from_f<T>(acc[d] * inv);
}

// ------------------------------------------------------------------ attention, few queries
constexpr int FQ_THREADS = 512, FQ_WARPS = 16, FQ_MAXQ = 16;

template <typename T, int D>
__global__ void __launch_bounds__(FQ_THREADS)
attn_fewq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                 long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts,
                 long o_bs, long o_hs, long o_ts, int Lq, int Lk, float scale) {
  MS2_PDL_WAIT();
  extern __shared__ float smf[];
  float* S = smf;                                   // [Lq][Lk]
  float* Qs = S + (long)Lq * Lk;                    // [Lq][D]
  float* red = Qs + FQ_MAXQ * D;                    // [FQ_WARPS][Lq][D]
  float* linv = red + FQ_WARPS * FQ_MAXQ * D;       // [Lq]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < Lq * D; idx += FQ_THREADS)
    Qs[idx] = to_f(q[b * q_bs + h * q_hs + (long)(idx / D) * q_ts + idx % D]) * scale;
  __syncthreads();
  // pass 1: scores
  const T* kb = k + b * k_bs + h * k_hs;
  for (int j = tid; j < Lk; j += FQ_THREADS) {
    float kv[D];
#pragma unroll
    for (int d = 0; d < D; d += 8) {
      float t[8];
      Vec8<T>::load(kb + (long)j * k_ts + d, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) kv[d + i] = t[i];
    }
    for (int qi = 0; qi < Lq; ++qi) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(Qs[qi * D + d], kv[d], a);
      S[(long)qi * Lk + j] = a;
    }
  }
  __syncthreads();
  // pass 2: softmax numerators per query row (one warp per row)
  for (int qi = warp; qi < Lq; qi += FQ_WARPS) {
    float* row = S + (long)qi * Lk;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float l = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float pj = __expf(row[j] - mx);
      row[j] = pj;
      l += pj;
    }
    l = warp_sum(l);
    if (lane == 0) linv[qi] = 1.f / l;
  }
  __syncthreads();
  // pass 3: O = P V ; lane -> (key slot within the warp's step, d)
  constexpr int KPW = 32 / D;                        // keys per warp step (2 for D=16, 1 for D=32)
  const int d = lane % D, sub = lane / D;
  float acc[FQ_MAXQ];
#pragma unroll
  for (int qi = 0; qi < FQ_MAXQ; ++qi) acc[qi] = 0.f;
  const T* vb = v + b * v_bs + h * v_hs;
  for (int j = warp * KPW + sub; j < Lk; j += FQ_WARPS * KPW) {
    const float vv = to_f(vb[(long)j * v_ts + d]);
#pragma unroll
    for (int qi = 0; qi < FQ_MAXQ; ++qi)
      if (qi < Lq) acc[qi] = fmaf(S[(long)qi * Lk + j], vv, acc[qi]);
  }
#pragma unroll
  for (int qi = 0; qi < FQ_MAXQ; ++qi) {
    float a = acc[qi];
    if (KPW == 2) a += __shfl_xor_sync(0xffffffffu, a, 16);
    if (qi < Lq && sub == 0) red[(warp * FQ_MAXQ + qi) * D + d] = a;
  }
  __syncthreads();
  for (int idx = tid; idx < Lq * D; idx += FQ_THREADS) {
    const int qi = idx / D, dd = idx % D;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < FQ_WARPS; ++w) a += red[(w * FQ_MAXQ + qi) * D + dd];
    o[b * o_bs + h * o_hs + (long)qi * o_ts + dd] = from_f<T>(a * linv[qi]);
  }
}

// ---- few queries, many keys, split over key chunks (flash-decoding layout): grid (Hh, B, NS) x 128 threads;
//      each CTA reduces its key chunk to (m, l, o[D]) per query, attn_fewq_combine_kernel merges the NS partials.
constexpr int FS_THREADS = 128, FS_MAXCHUNK = 128;

template <typename T, int D>
__global__ void __launch_bounds__(FS_THREADS)
attn_fewq_split_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, float* __restrict__ part,
                       long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts,
                       int Lq, int Lk, int chunk, float scale) {
  MS2_PDL_WAIT();
  __shared__ float S[FQ_MAXQ][FS_MAXCHUNK];
  __shared__ float Vs[FS_MAXCHUNK][D + 1];
  __shared__ float Qs[FQ_MAXQ][D];
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, ns = gridDim.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = sp * chunk;
  const int nk = min(chunk, Lk - j0);                      // >= 1 by construction
  for (int idx = tid; idx < Lq * D; idx += FS_THREADS)
    Qs[idx / D][idx % D] = to_f(q[b * q_bs + h * q_hs + (long)(idx / D) * q_ts + idx % D]) * scale;
  __syncthreads();
  const T* kb = k + b * k_bs + h * k_hs;
  const T* vb = v + b * v_bs + h * v_hs;
  for (int j = tid; j < nk; j += FS_THREADS) {
    float kv[D], vv[D];                                      // K and V rows of this key: all loads issued up front
#pragma unroll
    for (int d = 0; d < D; d += 8) {
      float t[8], u[8];
      Vec8<T>::load(kb + (long)(j0 + j) * k_ts + d, t);
      Vec8<T>::load(vb + (long)(j0 + j) * v_ts + d, u);
#pragma unroll
      for (int i = 0; i < 8; ++i) { kv[d + i] = t[i]; vv[d + i] = u[i]; }
    }
    for (int qi = 0; qi < Lq; ++qi) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(Qs[qi][d], kv[d], a);
      S[qi][j] = a;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) Vs[j][d] = vv[d];
  }
  __syncthreads();
  float* pbase = part + ((((long)b * gridDim.x + h) * ns + sp) * Lq) * (D + 2);
  for (int qi = warp; qi < Lq; qi += FS_THREADS / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, S[qi][j]);
    mx = warp_max(mx);
    float l = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float pj = __expf(S[qi][j] - mx);
      S[qi][j] = pj;
      l += pj;
    }
    l = warp_sum(l);
    if (lane == 0) {
      pbase[qi * (D + 2) + D] = mx;
      pbase[qi * (D + 2) + D + 1] = l;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < Lq * D; idx += FS_THREADS) {
    const int qi = idx / D, d = idx % D;
    float a = 0.f;
    for (int j = 0; j < nk; ++j) a = fmaf(S[qi][j], Vs[j][d], a);
    pbase[qi * (D + 2) + d] = a;
  }
}

// one warp per (batch*head, query): lane = key split (two rounds when ns > 32), so the D+2 loads of a lane are all
// independent and the merge is a handful of warp reductions (the per-thread loop over the splits walked the
// partials with one L2 round trip per unrolled batch: 8 us for 9 queries)
template <typename T, int D>
__global__ void __launch_bounds__(256)
attn_fewq_combine_kernel(const float* __restrict__ part, T* __restrict__ o, long o_bs, long o_hs, long o_ts,
                         int Hh, int Lq, int ns, int nrows) {
  MS2_PDL_WAIT();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int bh = row / Lq, qi = row - bh * Lq;
  const int b = bh / Hh, h = bh - b * Hh;
  const long sstride = (long)Lq * (D + 2);
  const float* pb = part + ((long)bh * ns * Lq + qi) * (D + 2);
  float m[2], l[2], v[2][D];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int s = lane + 32 * r;
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) v[r][d] = 0.f;
    if (s < ns) {
      const float* ps = pb + s * sstride;
      m[r] = ps[D];
      l[r] = ps[D + 1];
#pragma unroll
      for (int d = 0; d < D; ++d) v[r][d] = ps[d];
    }
  }
  const float mstar = warp_max(fmaxf(m[0], m[1]));
  const float w0 = __expf(m[0] - mstar), w1 = __expf(m[1] - mstar);       // exp(-inf) = 0 for absent splits
  const float lsum = warp_sum(w0 * l[0] + w1 * l[1]);
  const float inv = 1.f / lsum;
  T* op = o + b * o_bs + h * o_hs + (long)qi * o_ts;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float a = warp_sum(w0 * v[0][d] + w1 * v[1][d]);
    if (lane == (d & 31)) op[d] = from_f<T>(a * inv);
  }
}

size_t fewq_smem(int Lq, int Lk, int D) {
  return sizeof(float) * ((size_t)Lq * Lk + FQ_MAXQ * D + FQ_WARPS * FQ_MAXQ * D + FQ_MAXQ);
}

}  // namespace

// -------- dispatch helpers used by ms2_gemm / ms2_attention_ws
bool ms2_gemm_smallm_supported(int a_dt, int w_dt, const void* A, const void* W, long lda, int M, int N, int K) {
  if (a_dt != w_dt || M > 64 || K > 2048) return false;     // <= 8 register-resident weight chunks per lane
  const int vb = a_dt == MS2_BF16 ? 8 : 4;          // 16-byte vectors
  return K % 8 == 0 && lda % vb == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0);
}

int ms2_gemm_smallm_launch(const void* A, int a_dt, long lda, const void* W, const float* bias, const float* colscale,
                           const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K, int act,
                           cudaStream_t st) {
  const bool wide = K > 512;                         // one column per warp, 32 lanes split K (<= 8 chunks per lane)
  const int grid = ceil_div(N, SMK_WARPS * (wide ? 1 : 4));
  const int esz = a_dt == MS2_BF16 ? 2 : 4;
  int rpp = (int)((48 * 1024) / ((long)K * esz));    // rows of A staged per pass (<= 48 KB of shared memory)
  if (rpp > SM_MT) rpp = SM_MT;
  MS2_CHECK_ARG(rpp >= 1, "gemm_smallm: K = %d is too wide for the small-M kernel", K);
  const size_t smem = (size_t)(rpp < M ? rpp : M) * K * esz;
#define MS2_SMALLM(TA, TO)                                                                                          \
  do {                                                                                                              \
    if (wide)                                                                                                       \
      ms2_launch(gemm_smallm_kernel<TA, TO, 32, 8>, grid, SMK_WARPS * 32, smem, st, (const TA*)A, lda, (const TA*)W, bias, colscale, \
                                                                         residual, ldr, (TO*)out, ldo, M, N, K, act, rpp);  \
    else if (K <= 256)                                                                                              \
      ms2_launch(gemm_smallm_kernel<TA, TO, 8, 4>, grid, SMK_WARPS * 32, smem, st, (const TA*)A, lda, (const TA*)W, bias, colscale, \
                                                                        residual, ldr, (TO*)out, ldo, M, N, K, act, rpp);   \
    else                                                                                                            \
      ms2_launch(gemm_smallm_kernel<TA, TO, 8, 8>, grid, SMK_WARPS * 32, smem, st, (const TA*)A, lda, (const TA*)W, bias, colscale, \
                                                                        residual, ldr, (TO*)out, ldo, M, N, K, act, rpp);   \
  } while (0)
  if (a_dt == MS2_BF16 && o_dt == MS2_BF16) MS2_SMALLM(bf16, bf16);
  else if (a_dt == MS2_BF16 && o_dt == MS2_F32) MS2_SMALLM(bf16, float);
  else if (a_dt == MS2_F32 && o_dt == MS2_BF16) MS2_SMALLM(float, bf16);
  else if (a_dt == MS2_F32 && o_dt == MS2_F32) MS2_SMALLM(float, float);
  else {
    ms2_set_error("gemm_smallm: bad dtype");
    return MS2_ERR_ARG;
  }
#undef MS2_SMALLM
  MS2_CHECK_LAUNCH("gemm_smallm_kernel");
  return MS2_OK;
}

extern "C" int ms2_gemm_smallm_grouped(int groups, const void* const* h_A, const long* h_lda, const void* const* h_W,
                                       const float* const* h_bias, void* const* h_out, const long* h_ldo,
                                       const int* h_N, const int* h_act, int a_dt, int o_dt, int M, int K,
                                       void* stream) {
  MS2_CHECK_ARG(groups >= 1 && groups <= SM_MAXG, "gemm_smallm_grouped: 1..%d groups", SM_MAXG);
  MS2_CHECK_ARG(M >= 1 && M <= 64 && K > 0 && K % 8 == 0 && K <= 512, "gemm_smallm_grouped: M in 1..64, K %% 8 == 0, K <= 512");
  int nmax = 0;
  const int vb = a_dt == MS2_BF16 ? 8 : 4;
  for (int g = 0; g < groups; ++g) {
    MS2_CHECK_ARG(h_A[g] && h_W[g] && h_out[g] && h_N[g] > 0, "gemm_smallm_grouped: bad group %d", g);
    MS2_CHECK_ARG(h_lda[g] % vb == 0 && ((uintptr_t)h_A[g] % 16 == 0) && ((uintptr_t)h_W[g] % 16 == 0),
                  "gemm_smallm_grouped: group %d operands must be 16-byte aligned", g);
    nmax = h_N[g] > nmax ? h_N[g] : nmax;
  }
  dim3 grid(ceil_div(nmax, SM_WARPS * 4), groups);
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = a_dt == MS2_BF16 ? 2 : 4;
  int rpp = (int)((48 * 1024) / ((long)K * esz));
  if (rpp > SM_MT) rpp = SM_MT;
  MS2_CHECK_ARG(rpp >= 1, "gemm_smallm_grouped: K = %d is too wide", K);
  const size_t smem = (size_t)(rpp < M ? rpp : M) * K * esz;
#define MS2_GROUPED(TA, TO)                                                                        \
  do {                                                                                             \
    GroupedP<TA> p;                                                                                \
    memset(&p, 0, sizeof(p));                                                                      \
    for (int g = 0; g < groups; ++g) {                                                             \
      p.A[g] = (const TA*)h_A[g]; p.W[g] = (const TA*)h_W[g]; p.bias[g] = h_bias ? h_bias[g] : nullptr; \
      p.out[g] = h_out[g]; p.lda[g] = h_lda[g]; p.ldo[g] = h_ldo[g]; p.N[g] = h_N[g];              \
      p.act[g] = h_act ? h_act[g] : 0;                                                             \
    }                                                                                              \
    if (K <= 256) ms2_launch(gemm_smallm_grouped_kernel<TA, TO, 4>, grid, SM_WARPS * 32, smem, st, p, M, K, rpp);         \
    else ms2_launch(gemm_smallm_grouped_kernel<TA, TO, 8>, grid, SM_WARPS * 32, smem, st, p, M, K, rpp);                   \
  } while (0)
  if (a_dt == MS2_BF16 && o_dt == MS2_BF16) MS2_GROUPED(bf16, bf16);
  else if (a_dt == MS2_BF16 && o_dt == MS2_F32) MS2_GROUPED(bf16, float);
  else if (a_dt == MS2_F32 && o_dt == MS2_BF16) MS2_GROUPED(float, bf16);
  else if (a_dt == MS2_F32 && o_dt == MS2_F32) MS2_GROUPED(float, float);
  else {
    ms2_set_error("gemm_smallm_grouped: bad dtype");
    return MS2_ERR_ARG;
  }
#undef MS2_GROUPED
  MS2_CHECK_LAUNCH("gemm_smallm_grouped_kernel");
  return MS2_OK;
}

// returns 1 if a small-shape kernel handled the call, 0 if not applicable, <0 on error
int ms2_attention_small(const void* q, const void* k, const void* v, void* o, int dt, long q_bs, long q_hs, long q_ts,
                        long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs,
                        long o_ts, int B, int Hh, int Lq, int Lk, int D, float scale, void* ws, long ws_bytes,
                        cudaStream_t st) {
  if (!(D == 16 || D == 32)) return 0;
  const int vb = dt == MS2_BF16 ? 8 : 4;
  const bool vec_ok = ((q_ts | k_ts | q_hs | k_hs | q_bs | k_bs | v_ts | v_hs | v_bs) % vb == 0) && ((uintptr_t)q % 16 == 0) &&
                      ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0);
  if (!vec_ok) return 0;
#define MS2_ARGS_T(T) (const T*)q, (const T*)k, (const T*)v, (T*)o, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts
  if (Lk <= 32 && (size_t)2 * Hh * (Lk * D + 4) * 4 <= 48 * 1024) {
    const size_t smem = (size_t)2 * Hh * (Lk * D + 4) * 4;
    dim3 grid(ceil_div((long)Lq * Hh, 256), B);
    if (dt == MS2_BF16 && D == 16) ms2_launch(attn_fewk_kernel<bf16, 16>, grid, 256, smem, st, MS2_ARGS_T(bf16), Hh, Lq, Lk, scale);
    else if (dt == MS2_BF16 && D == 32) ms2_launch(attn_fewk_kernel<bf16, 32>, grid, 256, smem, st, MS2_ARGS_T(bf16), Hh, Lq, Lk, scale);
    else if (dt == MS2_F32 && D == 16) ms2_launch(attn_fewk_kernel<float, 16>, grid, 256, smem, st, MS2_ARGS_T(float), Hh, Lq, Lk, scale);
    else if (dt == MS2_F32 && D == 32) ms2_launch(attn_fewk_kernel<float, 32>, grid, 256, smem, st, MS2_ARGS_T(float), Hh, Lq, Lk, scale);
    else return 0;
    MS2_CHECK_LAUNCH("attn_fewk_kernel");
    return 1;
  }
  if (Lq <= FQ_MAXQ && Lk >= 512 && ws) {
    const int chunk = FS_MAXCHUNK;
    const int ns = (Lk + chunk - 1) / chunk;
    const long need = (long)B * Hh * ns * Lq * (D + 2) * 4;
    if (ns <= 64 && need <= ws_bytes) {
      dim3 grid(Hh, B, ns);
#define MS2_FEWQS(T, DD)                                                                                              \
  do {                                                                                                                \
    ms2_launch(attn_fewq_split_kernel<T, DD>, grid, FS_THREADS, 0, st, (const T*)q, (const T*)k, (const T*)v, (float*)ws, q_bs, \
                                                               q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, Lq, Lk, chunk, scale); \
    ms2_launch(attn_fewq_combine_kernel<T, DD>, ceil_div((long)B * Hh * Lq, 8), 256, 0, st, (const float*)ws, (T*)o, o_bs, o_hs, o_ts, Hh, Lq, ns, B * Hh * Lq); \
  } while (0)
      if (dt == MS2_BF16 && D == 16) MS2_FEWQS(bf16, 16);
      else if (dt == MS2_BF16 && D == 32) MS2_FEWQS(bf16, 32);
      else if (dt == MS2_F32 && D == 16) MS2_FEWQS(float, 16);
      else if (dt == MS2_F32 && D == 32) MS2_FEWQS(float, 32);
      else return 0;
#undef MS2_FEWQS
      MS2_CHECK_LAUNCH("attn_fewq_split_kernel");
      return 1;
    }
  }
  if (Lq <= FQ_MAXQ && fewq_smem(Lq, Lk, D) <= 200 * 1024) {
    const size_t smem = fewq_smem(Lq, Lk, D);
    dim3 grid(Hh, B);
#define MS2_FEWQ(T, DD)                                                                                         \
  do {                                                                                                          \
    auto kern = attn_fewq_kernel<T, DD>;                                                                        \
    static size_t attr = 0;                                                                                     \
    if (smem > attr) {                                                                                          \
      MS2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "fewq attr"); \
      attr = 200 * 1024;                                                                                        \
    }                                                                                                           \
    ms2_launch(kern, grid, FQ_THREADS, smem, st, MS2_ARGS_T(T), Lq, Lk, scale);                                         \
  } while (0)
    if (dt == MS2_BF16 && D == 16) MS2_FEWQ(bf16, 16);
    else if (dt == MS2_BF16 && D == 32) MS2_FEWQ(bf16, 32);
    else if (dt == MS2_F32 && D == 16) MS2_FEWQ(float, 16);
    else if (dt == MS2_F32 && D == 32) MS2_FEWQ(float, 32);
    else return 0;
#undef MS2_FEWQ
    MS2_CHECK_LAUNCH("attn_fewq_kernel");
    return 1;
  }
#undef MS2_ARGS_T
  return 0;
}
