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
  MS2_PDL_WAIT();
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

// Vectorised variant for C % 4 == 0, C <= 1024 (every LayerNorm of the encoder / memory attention / decoder image
// side): one warp per row, the row is read from HBM ONCE as float4 into registers (VPT float4 per lane), statistics by
// shuffles, 8- or 16-byte stores.  Same two-pass formula as above, so results are bit-identical to it.
template <typename TO, int VPT>
__global__ void __launch_bounds__(256)
layernorm_vec_kernel(const float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ gamma,
                     const float* __restrict__ beta, TO* __restrict__ y, long M, int C, float eps, int act) {
  MS2_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const int nv = C >> 2;
  const float4* xr = (const float4*)(x + row * (long)C);
  const float4* ar = add ? (const float4*)(add + row * (long)C) : nullptr;
  float4 v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nv) {
      v[i] = xr[c];
      if (ar) { const float4 a = ar[c]; v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w; }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (lane + i * 32 < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c2 * c2 + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  TO* yr = y + row * (long)C;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg((const float4*)gamma + c), bb = __ldg((const float4*)beta + c);
      float o[4] = {(v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                    (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (act == 1) o[j] = gelu_erf(o[j]);
        else if (act == 2) o[j] = fmaxf(o[j], 0.f);
      }
      if (sizeof(TO) == 4) {
        *(float4*)((float*)yr + c * 4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
        uint2 u;
        u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1;
        *(uint2*)((bf16*)yr + c * 4) = u;
      }
    }
  }
}

// Narrow rows (C <= 256): ROWS rows per warp, every row's loads issued before the first reduction - with one row per
// warp a thread has a single 16-byte load in flight and the stage-1 LayerNorm (524 288 rows x 96) ran at 3.2 TB/s.
// Same arithmetic per row as layernorm_vec_kernel (bit-identical results).
template <typename TO, int VPT, int ROWS>
__global__ void __launch_bounds__(256)
layernorm_vec_rows_kernel(const float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ gamma,
                          const float* __restrict__ beta, TO* __restrict__ y, long M, int C, float eps, int act) {
  MS2_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row0 = (((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ROWS;
  if (row0 >= M) return;
  const int nv = C >> 2;
  float4 v[ROWS][VPT];
  float s[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long row = row0 + r < M ? row0 + r : M - 1;
    const float4* xr = (const float4*)(x + row * (long)C);
    const float4* ar = add ? (const float4*)(add + row * (long)C) : nullptr;
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + i * 32;
      v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nv) {
        v[r][i] = xr[c];
        if (ar) { const float4 a = ar[c]; v[r][i].x += a.x; v[r][i].y += a.y; v[r][i].z += a.z; v[r][i].w += a.w; }
        s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
      }
    }
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) mean[r] = warp_sum(s[r]) / (float)C;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (lane + i * 32 < nv) {
        const float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], c2 = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
        q += (a * a + b * b) + (c2 * c2 + d * d);
      }
    }
    rstd[r] = rsqrtf(warp_sum(q) / (float)C + eps);
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg((const float4*)gamma + c), bb = __ldg((const float4*)beta + c);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (row0 + r >= M) continue;
        TO* yr = y + (row0 + r) * (long)C;
        float o[4] = {(v[r][i].x - mean[r]) * rstd[r] * g.x + bb.x, (v[r][i].y - mean[r]) * rstd[r] * g.y + bb.y,
                      (v[r][i].z - mean[r]) * rstd[r] * g.z + bb.z, (v[r][i].w - mean[r]) * rstd[r] * g.w + bb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (act == 1) o[j] = gelu_erf(o[j]);
          else if (act == 2) o[j] = fmaxf(o[j], 0.f);
        }
        if (sizeof(TO) == 4) {
          *(float4*)((float*)yr + c * 4) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
          uint2 u;
          u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1;
          *(uint2*)((bf16*)yr + c * 4) = u;
        }
      }
    }
  }
}

template <typename TO>
int launch_ln(const float* x, const float* add, const float* gamma, const float* beta, TO* y, long M, int C,
              float eps, int act, cudaStream_t st) {
  const int threads = 256;
  const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!add || (uintptr_t)add % 16 == 0) &&
                       ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0);
  if (C % 4 == 0 && C >= 32 && C <= 1024 && aligned) {
    const int vpt = (C / 4 + 31) / 32;
    const int grid = ceil_div(M * 32, threads);
#define LN_VEC(V) ms2_launch(layernorm_vec_kernel<TO, V>, grid, threads, 0, st, x, add, gamma, beta, y, M, C, eps, act)
    if (vpt == 1 && M >= 4096) {
      ms2_launch(layernorm_vec_rows_kernel<TO, 1, 4>, ceil_div(ceil_div(M, 4) * 32, threads), threads, 0, st, x, add, gamma, beta, y, M,
                                                                                                  C, eps, act);
    } else if (vpt == 2 && M >= 4096) {
      ms2_launch(layernorm_vec_rows_kernel<TO, 2, 2>, ceil_div(ceil_div(M, 2) * 32, threads), threads, 0, st, x, add, gamma, beta, y, M,
                                                                                                  C, eps, act);
    } else if (vpt == 1) LN_VEC(1);
    else if (vpt == 2) LN_VEC(2);
    else if (vpt <= 4) LN_VEC(4);
    else LN_VEC(8);
#undef LN_VEC
    MS2_CHECK_LAUNCH("layernorm_vec_kernel");
    return MS2_OK;
  }
#define LN_LAUNCH(G)                                                                            \
  ms2_launch(layernorm_kernel<TO, G>, ceil_div(M * G, threads), threads, 0, st, x, add, gamma, beta, y, M, C, eps, act)
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
