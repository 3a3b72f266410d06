// Shared helpers for the medsam2_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MS2_OK 0
#define MS2_ERR_ARG (-1)
#define MS2_ERR_CUDA (-2)
#define MS2_ERR_UNSUPPORTED (-3)

#define MS2_F32 0
#define MS2_BF16 1

extern "C" void ms2_set_error(const char* fmt, ...);

#define MS2_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      ms2_set_error(__VA_ARGS__);                      \
      return MS2_ERR_ARG;                              \
    }                                                  \
  } while (0)

#define MS2_CHECK_LAUNCH(name)                                             \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      ms2_set_error("%s: %s", name, cudaGetErrorString(e__));              \
      return MS2_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

#define MS2_CUDA(call, name)                                               \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      ms2_set_error("%s: %s", name, cudaGetErrorString(e__));              \
      return MS2_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL).  The tracked-frame path is a chain of ~200 dependent kernels of 2-20 us each;
// with plain stream order kernel N+1 is only scheduled after kernel N has drained.  Every kernel of this library is
// launched with programmaticStreamSerialization and starts with MS2_PDL_WAIT(): `griddepcontrol.wait` blocks until the
// preceding kernel of the stream has COMPLETED and its writes are visible (so no kernel ever reads or overwrites data the
// previous one still uses), `griddepcontrol.launch_dependents` then lets the NEXT kernel be scheduled right away: its
// launch latency, block scheduling and (where the wait sits behind it) barrier / tensor-memory set-up overlap this
// kernel's execution.  Without the launch attribute (MS2_PDL=0) both instructions are no-ops.
#define MS2_PDL_WAIT() asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory")

#include <stdlib.h>
#include <utility>
static inline bool ms2_pdl_enabled() {
  static const bool on = []() { const char* e = getenv("MS2_PDL"); return !e || atoi(e) != 0; }();
  return on;
}
// cluster_x > 1: launch as thread-block clusters of cluster_x CTAs along x (kernels without __cluster_dims__)
template <typename... KArgs, typename... Args>
static inline cudaError_t ms2_launch_cluster(void (*kern)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (ms2_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t ms2_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  return ms2_launch_cluster(kern, 1, grid, block, smem, st, std::forward<Args>(args)...);
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU with erf by Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): one MUFU.RCP + one MUFU.EX2 instead of
// the ~25-instruction erff; used where the result is rounded to bf16 anyway (tensor-core GEMM epilogues).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);          // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// GELU for bf16 outputs: erf(z) ~= tanh(z * P(z^2)) (fit error 1.3e-5 on |z| <= 4, clamped beyond) with the
// single-instruction MUFU.TANH (rel. error 2^-11): total error is < 1/8 of the bf16 rounding of the result.
__device__ __forceinline__ float gelu_erf_tanh(float x) {
  float z = x * 0.70710678118654752440f;
  z = fminf(fmaxf(z, -4.0f), 4.0f);
  const float z2 = z * z;
  float p = fmaf(-0.00015486571f, z2, -0.0010999786f);
  p = fmaf(p, z2, 0.10336954f);
  p = fmaf(p, z2, 1.1282882f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// bilinear blend with a fixed operation order and rounding (no compiler-chosen FMA contraction): every kernel that
// up-samples mask logits (resize.cu, postproc.cu) produces the same bits for the same pixel
__device__ __forceinline__ float ms2_bilerp(float v00, float v01, float v10, float v11, float lx, float ly) {
  const float ax = __fsub_rn(1.f, lx), ay = __fsub_rn(1.f, ly);
  const float top = __fmaf_rn(lx, v01, __fmul_rn(ax, v00));
  const float bot = __fmaf_rn(lx, v11, __fmul_rn(ax, v10));
  return __fmaf_rn(ly, bot, __fmul_rn(ay, top));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// dtype dispatch helper: calls F<T>() with T = float or bf16
#define MS2_DISPATCH_DTYPE(dt, T, ...)                                     \
  do {                                                                     \
    if ((dt) == MS2_F32) { typedef float T; __VA_ARGS__; }                 \
    else if ((dt) == MS2_BF16) { typedef bf16 T; __VA_ARGS__; }            \
    else { ms2_set_error("bad dtype %d", (int)(dt)); return MS2_ERR_ARG; } \
  } while (0)
