// LayerNorm over the last dimension (nn.LayerNorm in Hiera / memory attention / mask decoder and
// LayerNorm2d of the reference's NCHW convs, which are token-major here).  fp32 statistics, two-pass
// variance like torch; optional fused pre-add and activation; output fp32 or bf16 (GEMM operand).
// Memory-bound: 4*M*C bytes (fp32 in + fp32 out) or 6*M*C/... per SURVEY §8(d); a group of G lanes
// owns one row so that tiny-C rows (C=4,16 in the mask down-sampler) still fill the warp.
#include "common.cuh"

namespace {

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TO, int G>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ gamma,
                 const float* __restrict__ beta, TO* __restrict__ y, long M, int C, float eps, int act) {
  const int lane_in_group = threadIdx.x % G;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool active = row < M;
  const float* xr = x + (active ? row : 0) * (long)C;
  const float* ar = add ? add + (active ? row : 0) * (long)C : nullptr;
  float s = 0.f;
  if (active)
    for (int c = lane_in_group; c < C; c += G) s += xr[c] + (ar ? ar[c] : 0.f);
  const float mean = group_sum<G>(s) / (float)C;
  float v = 0.f;
  if (active)
    for (int c = lane_in_group; c < C; c += G) {
      float d = xr[c] + (ar ? ar[c] : 0.f) - mean;
      v += d * d;
    }
  const float rstd = rsqrtf(group_sum<G>(v) / (float)C + eps);
  if (!active) return;
  TO* yr = y + row * (long)C;
  for (int c = lane_in_group; c < C; c += G) {
    float o = (xr[c] + (ar ? ar[c] : 0.f) - mean) * rstd * gamma[c] + beta[c];
    if (act == 1) o = gelu_erf(o);
    else if (act == 2) o = fmaxf(o, 0.f);
    yr[c] = from_f<TO>(o);
  }
}

template <typename TO>
int launch_ln(const float* x, const float* add, const float* gamma, const float* beta, TO* y, long M, int C,
              float eps, int act, cudaStream_t st) {
  const int threads = 256;
#define LN_LAUNCH(G)                                                                            \
  layernorm_kernel<TO, G><<<ceil_div(M * G, threads), threads, 0, st>>>(x, add, gamma, beta, y, M, C, eps, act)
  if (C <= 4) LN_LAUNCH(4);
  else if (C <= 8) LN_LAUNCH(8);
  else if (C <= 16) LN_LAUNCH(16);
  else LN_LAUNCH(32);
#undef LN_LAUNCH
  MS2_CHECK_LAUNCH("layernorm_kernel");
  return MS2_OK;
}

}  // namespace

extern "C" int ms2_layernorm(const float* x, const float* add, const float* gamma, const float* beta, void* y,
                             int y_dt, int M, int C, float eps, int act, void* stream) {
  MS2_CHECK_ARG(x && gamma && beta && y, "layernorm: null pointer");
  MS2_CHECK_ARG(M >= 0 && C > 0, "layernorm: bad shape");
  if (M == 0) return MS2_OK;
  MS2_DISPATCH_DTYPE(y_dt, TO, return launch_ln<TO>(x, add, gamma, beta, (TO*)y, M, C, eps, act, (cudaStream_t)stream));
  return MS2_OK;
}
