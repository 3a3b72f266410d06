// Post-processing after the per-slice path, HBM-bound kernels:
//   ms2_non_overlap      — SAM2Base._apply_non_overlapping_constraints (modeling/sam2_base.py:812-830; callers
//                          sam2_video_predictor.py:742-743 and :850-851): per pixel only the first arg-max object keeps
//                          its score, the others are clamped to <= -10.  8*N*P bytes, one pass.
//   ms2_score_lowres     — the validation loop's scoring of one volume straight from the LOW-resolution logits
//                          (func_3d/function.py:283-305 + func_3d/utils.py:139-214): bilinear x(H/h) up-sampling
//                          (sam2_video_predictor.py:736), the thresholded |p&g|, |p|, |g| counts of `eval_seg` and the
//                          BCE-with-logits sum in ONE pass over the ground truth.  The fp32 video-resolution logits are
//                          never written or re-read: algorithmic bytes 4*N*H*W (ground truth) + 4*N*h*w instead of
//                          4*N*h*w + 4*N*H*W (resize write) + 2 * 8*N*H*W (two scoring passes).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ non-overlapping constraints
__global__ void __launch_bounds__(256) non_overlap_kernel(const float* __restrict__ in, float* __restrict__ out, int n_obj,
                                                          long P4) {
  MS2_PDL_WAIT();
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P4; i += stride) {
    float4 best = __ldg((const float4*)in + i);
    int4 arg = make_int4(0, 0, 0, 0);
    for (int o = 1; o < n_obj; ++o) {
      const float4 v = __ldg((const float4*)in + (long)o * P4 + i);
      if (v.x > best.x) { best.x = v.x; arg.x = o; }          // strict >: ties keep the FIRST object (torch.argmax)
      if (v.y > best.y) { best.y = v.y; arg.y = o; }
      if (v.z > best.z) { best.z = v.z; arg.z = o; }
      if (v.w > best.w) { best.w = v.w; arg.w = o; }
    }
    for (int o = 0; o < n_obj; ++o) {
      float4 v = __ldg((const float4*)in + (long)o * P4 + i);  // second read hits L1/L2 (same CTA, just touched)
      if (arg.x != o) v.x = fminf(v.x, -10.f);
      if (arg.y != o) v.y = fminf(v.y, -10.f);
      if (arg.z != o) v.z = fminf(v.z, -10.f);
      if (arg.w != o) v.w = fminf(v.w, -10.f);
      __stcs((float4*)out + (long)o * P4 + i, v);
    }
  }
}
__global__ void non_overlap_scalar_kernel(const float* __restrict__ in, float* __restrict__ out, int n_obj, long P) {
  MS2_PDL_WAIT();
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
    float best = in[i];
    int arg = 0;
    for (int o = 1; o < n_obj; ++o) {
      const float v = in[(long)o * P + i];
      if (v > best) { best = v; arg = o; }
    }
    for (int o = 0; o < n_obj; ++o) {
      const float v = in[(long)o * P + i];
      out[(long)o * P + i] = arg == o ? v : fminf(v, -10.f);
    }
  }
}

// ------------------------------------------------------------------ fused up-sample + eval_seg counts + BCE sum
struct ScoreThr { float v[8]; };

// One thread owns 4 consecutive output columns of one output row (one 16-byte ground-truth load); the two source rows
// of the low-resolution plane stay in L1/L2 (a 256x256 plane is 256 KB).  The interpolation uses exactly the expression
// of bilinear_vec4_kernel (resize.cu) with explicit round-to-nearest intrinsics in BOTH kernels, so the thresholded
// counts are bit-identical to resize-then-count.
template <int T, bool BCE>
__global__ void __launch_bounds__(256) score_lowres_kernel(const float* __restrict__ low, const float* __restrict__ gt,
                                                           ScoreThr thr, float pwm1, int32_t* __restrict__ counts,
                                                           double* __restrict__ sums, int h, int w, int H, int W,
                                                           float sh, float sw) {
  MS2_PDL_WAIT();
  __shared__ int shc[3 * T];
  __shared__ double shs[8];
  const int n = blockIdx.y;
  const float* lp = low + (long)n * h * w;
  const float* gp = gt + (long)n * H * W;
  float c[3 * T];
#pragma unroll
  for (int k = 0; k < 3 * T; ++k) c[k] = 0.f;
  if (threadIdx.x < 3 * T) shc[threadIdx.x] = 0;
  double bsum = 0.0;
  const unsigned W4 = W >> 2, n4 = (unsigned)H * W4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned xq = i % W4, yo = i / W4;
    const float4 g = __ldcs((const float4*)gp + i);
    const float sy = fmaxf(sh * (yo + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, y1 = y0 + (y0 < h - 1);
    const float ly = sy - y0;
    const float* p0 = lp + (long)y0 * w;
    const float* p1 = lp + (long)y1 * w;
    const float ge[4] = {g.x, g.y, g.z, g.w};
    float term4 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int xo = xq * 4 + e;
      const float sx = fmaxf(sw * (xo + 0.5f) - 0.5f, 0.f);
      const int x0 = (int)sx, x1 = x0 + (x0 < w - 1);
      const float lx = sx - x0;
      const float p = ms2_bilerp(__ldg(p0 + x0), __ldg(p0 + x1), __ldg(p1 + x0), __ldg(p1 + x1), lx, ly);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float fa = p > thr.v[t] ? 1.f : 0.f, fb = ge[e] > thr.v[t] ? 1.f : 0.f;
        c[3 * t] = fmaf(fa, fb, c[3 * t]);
        c[3 * t + 1] += fa;
        c[3 * t + 2] += fb;
      }
      if (BCE) term4 += (1.f - ge[e]) * p + (1.f + pwm1 * ge[e]) * (log1pf(expf(-fabsf(p))) + fmaxf(-p, 0.f));
    }
    if (BCE) bsum += (double)term4;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3 * T; ++k) {
    const int v = __reduce_add_sync(0xffffffffu, (int)c[k]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&shc[k], v);
  }
  if (BCE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if ((threadIdx.x & 31) == 0) shs[threadIdx.x >> 5] = bsum;
  }
  __syncthreads();
  if (threadIdx.x < 3 * T && shc[threadIdx.x]) atomicAdd(counts + (long)n * 3 * T + threadIdx.x, shc[threadIdx.x]);
  if (BCE && threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += shs[k];
    atomicAdd(sums + n, t);
  }
}

template <int T>
void score_launch(cudaStream_t st, const float* low, const float* gt, const ScoreThr& thr, float pwm1, int32_t* counts,
                  double* sums, int N, int h, int w, int H, int W) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long per_plane = (long)sms * 4 / N, by_size = ((long)H * (W / 4) + 2047) / 2048;
  if (per_plane > by_size) per_plane = by_size;
  if (per_plane < 1) per_plane = 1;
  dim3 grid((unsigned)per_plane, N);
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  if (sums) ms2_launch(score_lowres_kernel<T, true>, grid, 256, 0, st, low, gt, thr, pwm1, counts, sums, h, w, H, W, sh, sw);
  else ms2_launch(score_lowres_kernel<T, false>, grid, 256, 0, st, low, gt, thr, pwm1, counts, sums, h, w, H, W, sh, sw);
}

}  // namespace

extern "C" int ms2_non_overlap(const float* in, float* out, int n_obj, long P, void* stream) {
  MS2_CHECK_ARG(n_obj >= 0 && P >= 0, "non_overlap: bad sizes");
  if (!n_obj || !P) return MS2_OK;
  MS2_CHECK_ARG(in && out && in != out, "non_overlap: null or aliased planes");
  cudaStream_t st = (cudaStream_t)stream;
  if (P % 4 == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) {
    long b = (P / 4 + 255) / 256;
    if (b > 148L * 8) b = 148L * 8;
    ms2_launch(non_overlap_kernel, (int)b, 256, 0, st, in, out, n_obj, P / 4);
  } else {
    long b = (P + 255) / 256;
    if (b > 148L * 8) b = 148L * 8;
    ms2_launch(non_overlap_scalar_kernel, (int)b, 256, 0, st, in, out, n_obj, P);
  }
  MS2_CHECK_LAUNCH("non_overlap");
  return MS2_OK;
}

extern "C" int ms2_score_lowres(const float* low, const float* gt, const float* h_thr, int T, float pos_weight,
                                int32_t* counts, double* sums, int N, int h, int w, int H, int W, void* stream) {
  MS2_CHECK_ARG(h_thr && T >= 1 && T <= 8 && N >= 0, "score_lowres: 1..8 thresholds");
  if (!N) return MS2_OK;
  MS2_CHECK_ARG(low && gt && counts && h > 0 && w > 0 && H > 0 && W > 0, "score_lowres: null planes or empty sizes");
  MS2_CHECK_ARG(W % 4 == 0 && (uintptr_t)gt % 16 == 0 && N <= 65535 && (long)H * W < (1L << 31),
                "score_lowres: W must be a multiple of 4, gt 16-byte aligned, at most 65535 planes");
  cudaStream_t st = (cudaStream_t)stream;
  MS2_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * T * N, st), "score_lowres memset");
  if (sums) MS2_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * N, st), "score_lowres memset");
  ScoreThr thr;
  for (int t = 0; t < 8; ++t) thr.v[t] = t < T ? h_thr[t] : 0.f;
  const float pwm1 = pos_weight - 1.f;
  switch (T) {
    case 1: score_launch<1>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 2: score_launch<2>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 3: score_launch<3>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 4: score_launch<4>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 5: score_launch<5>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 6: score_launch<6>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    case 7: score_launch<7>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
    default: score_launch<8>(st, low, gt, thr, pwm1, counts, sums, N, h, w, H, W); break;
  }
  MS2_CHECK_LAUNCH("score_lowres");
  return MS2_OK;
}
