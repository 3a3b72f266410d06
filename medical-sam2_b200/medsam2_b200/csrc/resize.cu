// Bilinear resampling of fp32 planes, align_corners=False (torch F.interpolate semantics):
//   plain     — mask logits x4 up-sampling (sam2_base.py:368), video-resolution output
//               (sam2_video_predictor.py:736,831,844), image-predictor post-processing (transforms.py:98)
//   antialias — triangle-filter down-sampling of mask prompts (sam2_base.py:321-327,421-427)
// HBM-bound: reads H*W*4, writes Ho*Wo*4 bytes per plane; one thread per output pixel, x fastest.
#include "common.cuh"

namespace {

__global__ void bilinear_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int Ho,
                                int Wo, float sh, float sw) {
  MS2_PDL_WAIT();
  const long n = (long)N * Ho * Wo;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int xo = i % Wo;
    long t = i / Wo;
    int yo = t % Ho;
    long pl = t / Ho;
    float sy = fmaxf(sh * (yo + 0.5f) - 0.5f, 0.f);
    float sx = fmaxf(sw * (xo + 0.5f) - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < W - 1);
    float ly = sy - y0, lx = sx - x0;
    const float* p = x + pl * (long)H * W;
    float v00 = p[(long)y0 * W + x0], v01 = p[(long)y0 * W + x1];
    float v10 = p[(long)y1 * W + x0], v11 = p[(long)y1 * W + x1];
    y[i] = ms2_bilerp(v00, v01, v10, v11, lx, ly);
  }
}

// 4 consecutive output columns per thread (same arithmetic per element), 32-bit indices, one 16-byte store:
// Wo % 4 == 0 and < 2^31 output elements
__global__ void bilinear_vec4_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned n4, int H, int W,
                                     unsigned Ho, unsigned Wo4, float sh, float sw) {
  MS2_PDL_WAIT();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned xq = i % Wo4;
    unsigned t = i / Wo4;
    const unsigned yo = t % Ho;
    const unsigned pl = t / Ho;
    const float sy = fmaxf(sh * (yo + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, y1 = y0 + (y0 < H - 1);
    const float ly = sy - y0;
    const float* p0 = x + ((long)pl * H + y0) * W;
    const float* p1 = x + ((long)pl * H + y1) * W;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int xo = xq * 4 + e;
      const float sx = fmaxf(sw * (xo + 0.5f) - 0.5f, 0.f);
      const int x0 = (int)sx, x1 = x0 + (x0 < W - 1);
      const float lx = sx - x0;
      const float v00 = p0[x0], v01 = p0[x1], v10 = p1[x0], v11 = p1[x1];
      o[e] = ms2_bilerp(v00, v01, v10, v11, lx, ly);
    }
    *(float4*)(y + 4L * i) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ void aa_bounds(int o, float scale, int in, float& center, float& support, float& inv,
                                          int& lo, int& size) {
  support = (scale >= 1.f) ? scale : 1.f;
  inv = (scale >= 1.f) ? 1.f / scale : 1.f;
  center = scale * (o + 0.5f);
  lo = max((int)(center - support + 0.5f), 0);
  size = min((int)(center + support + 0.5f), in) - lo;
}

__global__ void bilinear_aa_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int Ho,
                                   int Wo, float sh, float sw) {
  MS2_PDL_WAIT();
  const long n = (long)N * Ho * Wo;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int xo = i % Wo;
    long t = i / Wo;
    int yo = t % Ho;
    long pl = t / Ho;
    float cy, supy, invy, cx, supx, invx;
    int ylo, ysz, xlo, xsz;
    aa_bounds(yo, sh, H, cy, supy, invy, ylo, ysz);
    aa_bounds(xo, sw, W, cx, supx, invx, xlo, xsz);
    float wxs = 0.f;
    for (int j = 0; j < xsz; ++j) wxs += fmaxf(0.f, 1.f - fabsf((j + xlo - cx + 0.5f) * invx));
    float wys = 0.f;
    for (int j = 0; j < ysz; ++j) wys += fmaxf(0.f, 1.f - fabsf((j + ylo - cy + 0.5f) * invy));
    const float* p = x + pl * (long)H * W;
    float acc = 0.f;
    for (int jy = 0; jy < ysz; ++jy) {
      float wy = fmaxf(0.f, 1.f - fabsf((jy + ylo - cy + 0.5f) * invy)) / wys;
      float row = 0.f;
      for (int jx = 0; jx < xsz; ++jx) {
        float wx = fmaxf(0.f, 1.f - fabsf((jx + xlo - cx + 0.5f) * invx)) / wxs;
        row = fmaf(wx, p[(long)(ylo + jy) * W + xlo + jx], row);
      }
      acc = fmaf(wy, row, acc);
    }
    y[i] = acc;
  }
}

}  // namespace

extern "C" int ms2_resize_bilinear(const float* x, float* y, int N, int H, int W, int Ho, int Wo, int antialias,
                                   void* stream) {
  MS2_CHECK_ARG(x && y && H > 0 && W > 0 && Ho > 0 && Wo > 0, "resize_bilinear: bad args");
  long n = (long)N * Ho * Wo;
  if (!n) return MS2_OK;
  long blocks = (n + 255) / 256;
  int g = (int)(blocks > 148L * 32 ? 148L * 32 : blocks);
  float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  if (antialias) ms2_launch(bilinear_aa_kernel, g, 256, 0, (cudaStream_t)stream, x, y, N, H, W, Ho, Wo, sh, sw);
  else if (Wo % 4 == 0 && (long)N * Ho * Wo < (1L << 31) && ((uintptr_t)y % 16 == 0)) {
    const long n4 = (long)N * Ho * Wo / 4;
    long b = (n4 + 255) / 256;
    if (b > 148L * 16) b = 148L * 16;
    ms2_launch(bilinear_vec4_kernel, (int)b, 256, 0, (cudaStream_t)stream, x, y, (unsigned)n4, H, W, (unsigned)Ho,
                                                                  (unsigned)(Wo / 4), sh, sw);
  } else ms2_launch(bilinear_kernel, g, 256, 0, (cudaStream_t)stream, x, y, N, H, W, Ho, Wo, sh, sw);
  MS2_CHECK_LAUNCH("resize_bilinear");
  return MS2_OK;
}
