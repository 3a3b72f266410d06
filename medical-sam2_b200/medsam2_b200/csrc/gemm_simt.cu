// Reference-precision GEMM with the fused epilogue of include/medsam2_b200.h (ms2_gemm, impl=1):
// fp32 accumulate on the CUDA cores, any M/N/K, fp32 or bf16 operands.  This is the exact ("fp32
// mode") path and the on-device cross-check for the tcgen05 kernel in gemm_tc.cu, which takes over
// all large bf16 shapes; it is not meant to reach the tensor roofline.
#include "common.cuh"

int ms2_gemm_tc_launch(const void* A, long lda, const void* W, const float* bias, const float* colscale,
                       const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K,
                       int act, cudaStream_t st);
bool ms2_gemm_tc_supported(int a_dt, int w_dt, long lda, long ldo, int M, int N, int K);
bool ms2_gemm_smallm_supported(int a_dt, int w_dt, const void* A, const void* W, long lda, int M, int N, int K);
int ms2_gemm_smallm_launch(const void* A, int a_dt, long lda, const void* W, const float* bias, const float* colscale,
                           const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K, int act,
                           cudaStream_t st);

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TA, typename TW, typename TO>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, long lda, const TW* __restrict__ W, const float* __restrict__ bias,
                 const float* __restrict__ colscale, const float* __restrict__ residual, long ldr,
                 TO* __restrict__ out, long ldo, int M, int N, int K, int act) {
  MS2_PDL_WAIT();
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int gm = m0 + lrow;
      const int gn = n0 + lrow;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gk = k0 + lk + i;
        float av = 0.f, wv = 0.f;
        if (gk < K) {
          if (gm < M) av = to_f(A[(long)gm * lda + gk]);
          if (gn < N) wv = to_f(W[(long)gn * K + gk]);
        }
        As[lk + i][lrow] = av;
        Ws[lk + i][lrow] = wv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = fmaxf(v, 0.f);
      else if (act == 3) v = 1.f / (1.f + __expf(-v));
      if (colscale) v *= colscale[gn];
      if (residual) v += residual[(long)gm * ldr + gn];
      out[(long)gm * ldo + gn] = from_f<TO>(v);
    }
  }
}

template <typename TI, typename TO>
int launch(const void* A, long lda, const void* W, const float* bias, const float* colscale, const float* residual,
           long ldr, void* out, long ldo, int M, int N, int K, int act, cudaStream_t st) {
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  ms2_launch(gemm_simt_kernel<TI, TI, TO>, grid, 256, 0, st, (const TI*)A, lda, (const TI*)W, bias, colscale, residual, ldr,
                                                     (TO*)out, ldo, M, N, K, act);
  MS2_CHECK_LAUNCH("gemm_simt_kernel");
  return MS2_OK;
}

}  // namespace

extern "C" int ms2_gemm(const void* A, int a_dt, long lda, const void* W, int w_dt, const float* bias,
                        const float* colscale, const float* residual, long ldr, void* out, int o_dt, long ldo,
                        int M, int N, int K, int act, int impl, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MS2_CHECK_ARG(A && W && out, "gemm: null pointer");
  MS2_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  MS2_CHECK_ARG(a_dt == w_dt, "gemm: A and W dtypes must match (%d vs %d)", a_dt, w_dt);
  MS2_CHECK_ARG(lda >= K && ldo >= N, "gemm: bad leading dimension");
  if (M == 0) return MS2_OK;
  const bool tc_ok = ms2_gemm_tc_supported(a_dt, w_dt, lda, ldo, M, N, K);
  if (impl == 2) MS2_CHECK_ARG(tc_ok, "gemm: tcgen05 path does not support this shape/dtype");
  if (impl == 2 || (impl == 0 && tc_ok))
    return ms2_gemm_tc_launch(A, lda, W, bias, colscale, residual, ldr, out, o_dt, ldo, M, N, K, act, st);
  const bool sm_ok = ms2_gemm_smallm_supported(a_dt, w_dt, A, W, lda, M, N, K);
  if (impl == 3) MS2_CHECK_ARG(sm_ok, "gemm: small-M path does not support this shape/dtype");
  if (impl == 3 || (impl == 0 && sm_ok))
    return ms2_gemm_smallm_launch(A, a_dt, lda, W, bias, colscale, residual, ldr, out, o_dt, ldo, M, N, K, act, st);
  if (a_dt == MS2_F32) {
    if (o_dt == MS2_F32) return launch<float, float>(A, lda, W, bias, colscale, residual, ldr, out, ldo, M, N, K, act, st);
    if (o_dt == MS2_BF16) return launch<float, bf16>(A, lda, W, bias, colscale, residual, ldr, out, ldo, M, N, K, act, st);
  } else if (a_dt == MS2_BF16) {
    if (o_dt == MS2_F32) return launch<bf16, float>(A, lda, W, bias, colscale, residual, ldr, out, ldo, M, N, K, act, st);
    if (o_dt == MS2_BF16) return launch<bf16, bf16>(A, lda, W, bias, colscale, residual, ldr, out, ldo, M, N, K, act, st);
  }
  ms2_set_error("gemm: bad dtype");
  return MS2_ERR_ARG;
}
