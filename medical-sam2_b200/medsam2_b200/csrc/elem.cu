// Memory-bound element-wise / layout kernels of the hot path (HBM roofline; all are single-pass,
// coalesced along the channel (innermost) dimension, grid-stride with a grid sized to the SM count).
#include "common.cuh"

namespace {

inline int grid_for(long n, int threads = 256) {
  long b = (n + threads - 1) / threads;
  const long cap = 148L * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename TO>
__global__ void axpby_kernel(const float* __restrict__ x, float a, const float* __restrict__ z, float b, float c,
                             TO* __restrict__ y, long n, long zn) {
  MS2_PDL_WAIT();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(a * x[i] + (z ? b * z[i % zn] : 0.f) + c);
}

// 4 elements per thread (16-byte loads, 8/16-byte stores), the broadcast index of z computed once per vector with
// 32-bit arithmetic; used when n, zn are multiples of 4, n < 2^31 and the pointers are 16-byte aligned
__device__ __forceinline__ void store4(float* y, long i, float4 v) { *(float4*)(y + i) = v; }
__device__ __forceinline__ void store4(bf16* y, long i, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *(uint32_t*)&lo;
  u.y = *(uint32_t*)&hi;
  *(uint2*)(y + i) = u;
}
template <typename TO>
__global__ void axpby_vec4_kernel(const float* __restrict__ x, float a, const float* __restrict__ z, float b, float c,
                                  TO* __restrict__ y, unsigned n4, unsigned zn4) {
  MS2_PDL_WAIT();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const float4 xv = *(const float4*)(x + 4L * i);
    float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (z) zv = *(const float4*)(z + 4L * (i % zn4));
    store4(y, 4L * i, make_float4(a * xv.x + b * zv.x + c, a * xv.y + b * zv.y + c, a * xv.z + b * zv.z + c,
                                  a * xv.w + b * zv.w + c));
  }
}
template <typename TO>
__global__ void cast_vec4_kernel(const float* __restrict__ x, TO* __restrict__ y, unsigned n4) {
  MS2_PDL_WAIT();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    store4(y, 4L * i, *(const float4*)(x + 4L * i));
}

__global__ void gate_rows_kernel(const float* __restrict__ x, const float* __restrict__ gate, float fill,
                                 float* __restrict__ y, long n, long P) {
  MS2_PDL_WAIT();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = gate[i / P] > 0.f ? x[i] : fill;
}

__global__ void select_plane_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                    float* __restrict__ y, int B, int M, long P) {
  MS2_PDL_WAIT();
  const long n = (long)B * P;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long b = i / P, p = i - b * P;
    int m = idx[b];
    m = m < 0 ? 0 : (m >= M ? M - 1 : m);
    y[i] = x[(b * M + m) * P + p];
  }
}

__global__ void add_rowvec_kernel(const float* __restrict__ x, const float* __restrict__ v, float s,
                                  float* __restrict__ y, long n, int C) {
  MS2_PDL_WAIT();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = x[i] + s * v[i % C];
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long n) {
  MS2_PDL_WAIT();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(to_f(x[i]));
}

__global__ void act_kernel(const float* __restrict__ x, float* __restrict__ y, long n, int act) {
  MS2_PDL_WAIT();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (act == 1) v = gelu_erf(v);
    else if (act == 2) v = fmaxf(v, 0.f);
    else if (act == 3) v = 1.f / (1.f + expf(-v));
    y[i] = v;
  }
}

__global__ void upsample2x_add_kernel(float* __restrict__ fine, const float* __restrict__ coarse, int B, int H,
                                      int W, int C) {
  MS2_PDL_WAIT();
  const long n = (long)B * H * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = i % C;
    long t = i / C;
    int x = t % W; t /= W;
    int y = t % H;
    int b = t / H;
    fine[i] += coarse[(((long)b * (H / 2) + y / 2) * (W / 2) + x / 2) * C + c];
  }
}

// tiled transpose of [R, Cc] -> [Cc, R] per batch (NHWC<->NCHW with R=H*W)
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ y, long R, int Cc) {
  MS2_PDL_WAIT();
  __shared__ float tile[32][33];
  const long b = blockIdx.z;
  const float* xb = x + b * R * Cc;
  float* yb = y + b * R * Cc;
  long r0 = (long)blockIdx.y * 32;
  int c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long r = r0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? xb[r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long r = r0 + threadIdx.x;
    if (r < R && c < Cc) yb[(long)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  MS2_PDL_WAIT();
  const int Ho = H / 2, Wo = W / 2;
  const long n = (long)B * Ho * Wo * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = i % C;
    long t = i / C;
    int xo = t % Wo; t /= Wo;
    int yo = t % Ho;
    int b = t / Ho;
    const float* p = x + (((long)b * H + 2 * yo) * W + 2 * xo) * C + c;
    y[i] = fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(long)W * C], p[(long)W * C + C]));
  }
}

// 4 channels per thread (16-byte loads / stores), 32-bit indices: C % 4 == 0 and < 2^31 input elements
__global__ void maxpool2x2_vec4_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned n4, unsigned H,
                                       unsigned W, unsigned C4) {
  MS2_PDL_WAIT();
  const unsigned Ho = H / 2, Wo = W / 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4;
    unsigned t = i / C4;
    const unsigned xo = t % Wo; t /= Wo;
    const unsigned yo = t % Ho;
    const unsigned b = t / Ho;
    const float4* p = (const float4*)x + ((b * H + 2 * yo) * W + 2 * xo) * C4 + c4;
    const float4 a = p[0], bq = p[C4], c = p[W * C4], d = p[W * C4 + C4];
    *((float4*)y + i) = make_float4(fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y)),
                                    fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w)));
  }
}

__global__ void pixel_shuffle_add_kernel(const float* __restrict__ g, const float* __restrict__ bias,
                                         const float* __restrict__ skip, float* __restrict__ out, int B, int H,
                                         int W, int C, int act) {
  MS2_PDL_WAIT();
  const long n = (long)B * 2 * H * 2 * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = i % C;
    long t = i / C;
    int xo = t % (2 * W); t /= (2 * W);
    int yo = t % (2 * H);
    int b = t / (2 * H);
    int dy = yo & 1, dx = xo & 1;
    float v = g[((((long)b * H + (yo >> 1)) * W + (xo >> 1)) * 4 + dy * 2 + dx) * C + c];
    if (bias) v += bias[c];
    if (skip) v += skip[i];
    if (act == 1) v = gelu_erf(v);
    out[i] = v;
  }
}

// the same, 4 channels per thread (16-byte loads / stores), 32-bit index arithmetic: C % 4 == 0, < 2^31 elements
__global__ void pixel_shuffle_add_vec4_kernel(const float* __restrict__ g, const float* __restrict__ bias,
                                              const float* __restrict__ skip, float* __restrict__ out, unsigned n4,
                                              unsigned H, unsigned W, unsigned C4, int act) {
  MS2_PDL_WAIT();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4;
    unsigned t = i / C4;
    const unsigned xo = t % (2 * W); t /= (2 * W);
    const unsigned yo = t % (2 * H);
    const unsigned b = t / (2 * H);
    const unsigned src = ((((b * H + (yo >> 1)) * W + (xo >> 1)) * 4 + (yo & 1) * 2 + (xo & 1)) * C4 + c4);
    float4 v = *(const float4*)(g + 4L * src);
    if (bias) {
      const float4 bb = *(const float4*)(bias + 4 * c4);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    if (skip) {
      const float4 sk = *(const float4*)(skip + 4L * i);
      v.x += sk.x; v.y += sk.y; v.z += sk.z; v.w += sk.w;
    }
    if (act == 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
    *(float4*)(out + 4L * i) = v;
  }
}

// one warp per pixel: lanes own channels, Mk dot products reduced by shuffles
__global__ void hyper_mask_kernel(const float* __restrict__ up, const float* __restrict__ hyper,
                                  float* __restrict__ masks, int B, int P, int C, int Mk) {
  MS2_PDL_WAIT();
  __shared__ float hs[8 * 64];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < Mk * C; i += blockDim.x) hs[i] = hyper[(long)b * Mk * C + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + warp; p < P; p += gridDim.x * wpb) {
    const float* u = up + ((long)b * P + p) * C;
    float u0 = lane < C ? u[lane] : 0.f;
    float u1 = lane + 32 < C ? u[lane + 32] : 0.f;
    for (int m = 0; m < Mk; ++m) {
      float s = u0 * (lane < C ? hs[m * C + lane] : 0.f) + u1 * (lane + 32 < C ? hs[m * C + lane + 32] : 0.f);
      s = warp_sum(s);
      if (lane == 0) masks[((long)b * Mk + m) * P + p] = s;
    }
  }
}

// C = 32, Mk = 4 (the shipped decoder): one THREAD per pixel - its 32 channels are one 128-byte line (8 independent
// 16-byte loads), the 4 x 32 hyper-network weights are broadcast reads from shared memory, and the four outputs of
// consecutive pixels are coalesced stores (the warp-per-pixel kernel spent its time in 20 shuffles and 4 scalar stores
// per pixel: 17 us for 8 MB)
__global__ void __launch_bounds__(128)
hyper_mask_c32m4_kernel(const float* __restrict__ up, const float* __restrict__ hyper, float* __restrict__ masks, int P) {
  MS2_PDL_WAIT();
  __shared__ float4 hs[4 * 8];
  const int b = blockIdx.y;
  if (threadIdx.x < 32) hs[threadIdx.x] = *(const float4*)(hyper + (long)b * 128 + threadIdx.x * 4);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4* u = (const float4*)(up + ((long)b * P + p) * 32);
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = u[i];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 h = hs[m * 8 + i];
      s = fmaf(x[i].x, h.x, s); s = fmaf(x[i].y, h.y, s); s = fmaf(x[i].z, h.z, s); s = fmaf(x[i].w, h.w, s);
    }
    masks[((long)b * 4 + m) * P + p] = s;
  }
}

__global__ void fourier_pe_kernel(const float* __restrict__ coords, const float* __restrict__ gauss,
                                  float* __restrict__ out, int n, int F) {
  MS2_PDL_WAIT();
  const long total = (long)n * F;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int f = i % F;
    long r = i / F;
    float cx = 2.f * coords[r * 2] - 1.f, cy = 2.f * coords[r * 2 + 1] - 1.f;
    float v = cx * gauss[f] + cy * gauss[F + f];
    v = 6.283185307179586f * v;
    out[r * 2 * F + f] = sinf(v);
    out[r * 2 * F + F + f] = cosf(v);
  }
}

// Sparse prompt embedding in ONE launch (prompt_encoder.py:79-101: +0.5 shift, optional padding point with label -1,
// normalisation by the image size, random-Fourier features, zeroing of "not a point" rows, label embedding row):
// out[b, j, :] = keep(label) * [sin | cos](2*pi * ((2*xy/size - 1) @ gauss)) + table[clamp(label + 1, 0, 4)]
// with label = -1 and xy = 0 for the padding point j == N.  Replaces ~14 element-wise torch kernels per slice.
__global__ void point_embed_kernel(const float* __restrict__ coords, const int* __restrict__ labels,
                                   const float* __restrict__ gauss, const float* __restrict__ table,
                                   float* __restrict__ out, int B, int N, int pad, int F, float inv_w, float inv_h,
                                   const float* __restrict__ prefix, int P) {
  MS2_PDL_WAIT();
  const int Np = N + (pad ? 1 : 0) + P;
  const long total = (long)B * Np * F;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long r = i / F;
    int j = (int)(r % Np);
    const int b = (int)(r / Np);
    if (j < P) {                    // constant rows in front of the prompt rows (the mask decoder's output tokens)
      out[r * 2 * F + f] = prefix[(long)j * 2 * F + f];
      out[r * 2 * F + F + f] = prefix[(long)j * 2 * F + F + f];
      continue;
    }
    j -= P;
    float x = 0.f, y = 0.f;
    int lab = -1;
    if (j < N) {
      x = coords[((long)b * N + j) * 2] + 0.5f;
      y = coords[((long)b * N + j) * 2 + 1] + 0.5f;
      lab = labels[(long)b * N + j];
    }
    const float cx = 2.f * (x * inv_w) - 1.f, cy = 2.f * (y * inv_h) - 1.f;
    const float v = 6.283185307179586f * (cx * gauss[f] + cy * gauss[F + f]);
    const float keep = lab != -1 ? 1.f : 0.f;
    int row = lab + 1;
    row = row < 0 ? 0 : (row > 4 ? 4 : row);
    const float* t = table + (long)row * 2 * F;
    float* o = out + r * 2 * F;
    o[f] = sinf(v) * keep + t[f];
    o[F + f] = cosf(v) * keep + t[F + f];
  }
}

// in_layout 0: fp32 NCHW (video tensor), 1: uint8 NHWC (decoded images), 2: uint8 NCHW (uint8 video tensor);
// OUT = float (the reference's frame dtype) or bf16 (what the patch-embed contraction rounds the frame to anyway under
// autocast: half the bytes written here and read there).  4 pixels of one row per thread.
template <typename OUT>
__global__ void __launch_bounds__(256) normalize_image_kernel(const void* __restrict__ x, int in_layout, OUT* __restrict__ out,
                                                              int B, int H, int W) {
  MS2_PDL_WAIT();
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const int W4 = (W + 3) >> 2;
  const long n = (long)B * 3 * H * W4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int xq, yh, c, b;
    if (n < (1L << 31)) {                        // 32-bit index arithmetic (64-bit divisions cost ~100 clk each)
      unsigned t = (unsigned)i;
      xq = t % (unsigned)W4; t /= (unsigned)W4;
      yh = t % (unsigned)H; t /= (unsigned)H;
      c = t % 3u;
      b = t / 3u;
    } else {
      long t = i;
      xq = t % W4; t /= W4;
      yh = t % H; t /= H;
      c = t % 3;
      b = t / 3;
    }
    const long o0 = (((long)b * 3 + c) * H + yh) * W + 4 * xq;
    const int cnt = min(4, W - 4 * xq);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (in_layout == 0) {
      const float* p = (const float*)x + o0;
      if (cnt == 4 && ((uintptr_t)p & 15) == 0) { const float4 q = __ldcs((const float4*)p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else for (int e = 0; e < cnt; ++e) v[e] = p[e];
    } else if (in_layout == 2) {
      const uint8_t* p = (const uint8_t*)x + o0;
      if (cnt == 4 && ((uintptr_t)p & 3) == 0) { const uchar4 q = *(const uchar4*)p; v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else for (int e = 0; e < cnt; ++e) v[e] = (float)p[e];
    } else {
      const uint8_t* p = (const uint8_t*)x + (((long)b * H + yh) * W + 4 * xq) * 3 + c;
      for (int e = 0; e < cnt; ++e) v[e] = (float)p[3 * e];
    }
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) r[e] = (v[e] / 255.0f - mean[c]) / stdv[c];
    OUT* po = out + o0;
    if (cnt == 4 && ((uintptr_t)po & (4 * sizeof(OUT) - 1)) == 0) {          // one 16-byte (fp32) / 8-byte (bf16) store
      if (sizeof(OUT) == 4) {
        *(float4*)po = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(r[0], r[1]), h1 = __floats2bfloat162_rn(r[2], r[3]);
        *(uint2*)po = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
      }
    } else {
      for (int e = 0; e < cnt; ++e) po[e] = from_f<OUT>(r[e]);
    }
  }
}

// Segmentation-metric counts (func_3d/utils.py:139-202 `eval_seg`, :204-214 `iou`, :215-240 `dice_coeff`): the reference
// binarises prediction and ground truth once per threshold, copies both to the host and reduces them there (for the
// default 5 thresholds: 10 full-resolution passes + 10 D2H copies per object and slice).  Here ONE pass over the two
// planes yields, for every threshold, |pred>th & gt>th|, |pred>th| and |gt>th| as exact integers; IoU and Dice
// follow on the host from 3*T integers per plane.  HBM roofline: 8 bytes per pixel, read once.
struct SegThr { float v[8]; };
template <int T, bool VEC>
__global__ void __launch_bounds__(256) seg_counts_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                         SegThr thr, int32_t* __restrict__ counts, long P) {
  MS2_PDL_WAIT();
  __shared__ int sh[3 * T];
  const int n = blockIdx.y;
  const float* pp = pred + (long)n * P;
  const float* gp = gt + (long)n * P;
  // 0/1 flags as floats: per pixel and threshold 2 FSET (ALU pipe) + 2 FADD + 1 FFMA (FMA pipe) instead of ~10 integer
  // instructions (the integer version was issue-bound at 61 % of HBM); a thread sees < 2^24 pixels, so the sums are exact
  float c[3 * T];
#pragma unroll
  for (int k = 0; k < 3 * T; ++k) c[k] = 0.f;
  if (threadIdx.x < 3 * T) sh[threadIdx.x] = 0;
  auto acc = [&](float a, float b) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float fa = a > thr.v[t] ? 1.f : 0.f, fb = b > thr.v[t] ? 1.f : 0.f;
      c[3 * t] = fmaf(fa, fb, c[3 * t]);
      c[3 * t + 1] += fa;
      c[3 * t + 2] += fb;
    }
  };
  auto acc4 = [&](const float4& a, const float4& b) { acc(a.x, b.x); acc(a.y, b.y); acc(a.z, b.z); acc(a.w, b.w); };
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const long P4 = P >> 2;
    const float4* p4 = (const float4*)pp;
    const float4* g4 = (const float4*)gp;
    for (; i + stride < P4; i += 2 * stride) {            // four 16-byte loads in flight per thread
      const float4 a0 = __ldcs(p4 + i), b0 = __ldcs(g4 + i), a1 = __ldcs(p4 + i + stride), b1 = __ldcs(g4 + i + stride);
      acc4(a0, b0);
      acc4(a1, b1);
    }
    if (i < P4) acc4(__ldcs(p4 + i), __ldcs(g4 + i));
  } else {
    for (; i < P; i += stride) acc(pp[i], gp[i]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3 * T; ++k) {
    const int v = __reduce_add_sync(0xffffffffu, (int)c[k]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 3 * T && sh[threadIdx.x]) atomicAdd(counts + (long)n * 3 * T + threadIdx.x, sh[threadIdx.x]);
}
// One wave: the grid never exceeds the resident CTA slots (a grid-stride row split over a partial second wave would
// double the time), at most one CTA per 8192 pixels of a plane.
template <int T, bool VEC>
void seg_counts_launch2(cudaStream_t st, const float* pred, const float* gt, const SegThr& thr, int32_t* counts, int N,
                        long P) {
  static int slots = 0;
  if (!slots) {
    int per_sm = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_counts_kernel<T, VEC>, 256, 0);
    slots = sms * (per_sm > 0 ? per_sm : 1);
  }
  long per_plane = slots / N, by_size = (P + 8191) / 8192;
  if (per_plane > by_size) per_plane = by_size;
  dim3 grid((unsigned)(per_plane < 1 ? 1 : per_plane), N);
  ms2_launch(seg_counts_kernel<T, VEC>, grid, 256, 0, st, pred, gt, thr, counts, P);
}
template <int T>
void seg_counts_launch(bool vec, cudaStream_t st, const float* pred, const float* gt, const SegThr& thr, int32_t* counts,
                       int N, long P) {
  if (vec) seg_counts_launch2<T, true>(st, pred, gt, thr, counts, N, P);
  else seg_counts_launch2<T, false>(st, pred, gt, thr, counts, N, P);
}

// Validation loss of func_3d/function.py:35-36,299 (`BCEWithLogitsLoss(pos_weight)`): per plane the SUM of
// (1-y)*x + (1 + (pw-1)*y) * (log1p(exp(-|x|)) + max(-x, 0)) in one pass (8 B/pixel); 4 pixels are added in fp32, the
// running sums are fp64 (per thread, per CTA and in the per-plane atomics), so the mean on the host is good to fp32 ulps.
template <bool VEC>
__global__ void __launch_bounds__(256) bce_logits_sum_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                             float pwm1, double* __restrict__ sums, long P) {
  MS2_PDL_WAIT();
  __shared__ double sh[8];
  const int n = blockIdx.y;
  const float* pp = pred + (long)n * P;
  const float* gp = gt + (long)n * P;
  auto term = [&](float x, float y) {
    return (1.f - y) * x + (1.f + pwm1 * y) * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
  };
  double acc = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const long P4 = P >> 2;
    for (; i < P4; i += stride) {
      const float4 a = __ldcs((const float4*)pp + i), b = __ldcs((const float4*)gp + i);
      acc += (double)((term(a.x, b.x) + term(a.y, b.y)) + (term(a.z, b.z) + term(a.w, b.w)));
    }
  } else {
    for (; i < P; i += stride) acc += (double)term(pp[i], gp[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(sums + n, t);
  }
}

// Automatic-mask-generator statistics (automatic_mask_generator.py:300-340 + utils/amg.py:158-180 `calculate_stability_score`,
// :296-348 `batched_mask_to_box`): the reference makes two thresholded int16/int32 reductions, one binarisation and six
// max/min reductions over every [H,W] logit plane; here ONE pass (4 B/pixel) yields per plane
// (#(x > thr+off), #(x > thr-off), #(x > thr), min col, min row, max col, max row of x > thr).
__global__ void mask_stats_init_kernel(int32_t* __restrict__ out, int N, int H, int W) {
  MS2_PDL_WAIT();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int32_t* o = out + (long)n * 7;
  o[0] = o[1] = o[2] = 0;
  o[3] = W; o[4] = H; o[5] = -1; o[6] = -1;
}
__global__ void __launch_bounds__(256) mask_stats_kernel(const float* __restrict__ x, int32_t* __restrict__ out, int H, int W,
                                                         float thr, float off, bool vec) {
  MS2_PDL_WAIT();
  const int n = blockIdx.y;
  const float* xp = x + (long)n * H * W;
  const float hi = thr + off, lo = thr - off;
  int chi = 0, clo = 0, cmid = 0, x0 = W, y0 = H, x1 = -1, y1 = -1;
  float fhi = 0.f, flo = 0.f, fmid = 0.f, colf[4] = {0.f, 0.f, 0.f, 0.f};
  // a CTA walks whole rows (coalesced along the row); rows are dealt round-robin over the CTAs of the plane
  if (vec && W <= 4 * (int)blockDim.x) {
    // fast path (16-byte loads; every thread owns the SAME four columns in every row, four rows in flight): counts as 0/1
    // floats (FSET + FADD, exact: a thread sees < 2^24 pixels), "column ever set" as four float sums, rows per float4
    const int c = 4 * threadIdx.x, g = gridDim.x;
    auto process = [&](const float4& v, int r) {
      const float e[4] = {v.x, v.y, v.z, v.w};
      float rowsum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        fhi += e[j] > hi ? 1.f : 0.f;
        flo += e[j] > lo ? 1.f : 0.f;
        const float m = e[j] > thr ? 1.f : 0.f;
        rowsum += m;
        colf[j] += m;
      }
      fmid += rowsum;
      if (rowsum > 0.f) { y0 = min(y0, r); y1 = max(y1, r); }
    };
    if (c < W) {
      int r = blockIdx.x;
      for (; r + 3 * g < H; r += 4 * g) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs((const float4*)(xp + (long)(r + u * g) * W + c));
#pragma unroll
        for (int u = 0; u < 4; ++u) process(v[u], r + u * g);
      }
      for (; r < H; r += g) process(__ldcs((const float4*)(xp + (long)r * W + c)), r);
    }
  } else {
    for (int r = blockIdx.x; r < H; r += gridDim.x) {
      const float* row = xp + (long)r * W;
      bool any = false;
      auto visit = [&](float v, int c) {
        chi += v > hi;
        clo += v > lo;
        if (v > thr) {
          ++cmid;
          any = true;
          x0 = min(x0, c);
          x1 = max(x1, c);
        }
      };
      if (vec) {                                         // 16-byte loads: W % 4 == 0 and a 16-byte aligned base
        for (int c = 4 * threadIdx.x; c < W; c += 4 * blockDim.x) {
          const float4 v = __ldcs((const float4*)(row + c));
          visit(v.x, c); visit(v.y, c + 1); visit(v.z, c + 2); visit(v.w, c + 3);
        }
      } else {
        for (int c = threadIdx.x; c < W; c += blockDim.x) visit(row[c], c);
      }
      if (any) { y0 = min(y0, r); y1 = max(y1, r); }
    }
  }
  chi += (int)fhi; clo += (int)flo; cmid += (int)fmid;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (colf[j] > 0.f) { x0 = min(x0, 4 * (int)threadIdx.x + j); x1 = max(x1, 4 * (int)threadIdx.x + j); }
  chi = __reduce_add_sync(0xffffffffu, chi);
  clo = __reduce_add_sync(0xffffffffu, clo);
  cmid = __reduce_add_sync(0xffffffffu, cmid);
  x0 = __reduce_min_sync(0xffffffffu, x0);
  y0 = __reduce_min_sync(0xffffffffu, y0);
  x1 = __reduce_max_sync(0xffffffffu, x1);
  y1 = __reduce_max_sync(0xffffffffu, y1);
  if ((threadIdx.x & 31) == 0) {
    int32_t* o = out + (long)n * 7;
    if (chi) atomicAdd(o, chi);
    if (clo) atomicAdd(o + 1, clo);
    if (cmid) {
      atomicAdd(o + 2, cmid);
      atomicMin(o + 3, x0); atomicMin(o + 4, y0); atomicMax(o + 5, x1); atomicMax(o + 6, y1);
    }
  }
}

// Binarised, un-cropped, TRANSPOSED copies of selected planes for run-length encoding (utils/amg.py:279-293 `uncrop_masks`,
// :107-134 `mask_to_rle_pytorch`, which encodes in column-major order): out[k, x0 + c, y0 + r] = x[sel[k], r, c] > thr as
// uint8, out is [K, OW, OH] and zero outside the crop.  32x32 shared-memory tiles: coalesced reads and writes.
__global__ void mask_binarize_t_kernel(const float* __restrict__ x, const int32_t* __restrict__ sel, uint8_t* __restrict__ out,
                                       int H, int W, float thr, int OH, int OW, int x0, int y0) {
  MS2_PDL_WAIT();
  __shared__ uint8_t tile[32][33];
  const int k = blockIdx.z;
  const float* xp = x + (long)sel[k] * H * W;
  uint8_t* op = out + (long)k * OH * OW;
  const int c = blockIdx.x * 32 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + threadIdx.y + j;
    tile[threadIdx.y + j][threadIdx.x] = (r < H && c < W) ? (uint8_t)(xp[(long)r * W + c] > thr) : (uint8_t)0;
  }
  __syncthreads();
  const int r = blockIdx.y * 32 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int cc = blockIdx.x * 32 + threadIdx.y + j;
    if (r < H && cc < W) op[(long)(x0 + cc) * OH + (y0 + r)] = tile[threadIdx.x][threadIdx.y + j];
  }
}

// Run-length encoding on the device (utils/amg.py:107-134 `mask_to_rle_pytorch` finds the change positions with a
// whole-tensor XOR + `nonzero` and slices them per mask on the host): one CTA per mask walks its bytes in order, 16 per thread
// and step, and appends the positions p with m[p] != m[p-1] through a block-wide exclusive scan, so pos[k, 0..cnt[k]) is
// sorted.  cnt[k] is the TOTAL number of transitions; positions beyond `cap` are dropped (the caller re-runs with a larger cap).
__global__ void __launch_bounds__(1024) rle_transitions_kernel(const uint8_t* __restrict__ m, long L, int32_t* __restrict__ pos,
                                                               int32_t* __restrict__ cnt, int cap, bool vec) {
  MS2_PDL_WAIT();
  __shared__ int warp_tot[32];
  __shared__ int base_sh;
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t* mp = m + (long)k * L;
  int32_t* out = pos + (long)k * cap;
  if (tid == 0) base_sh = 0;
  __syncthreads();
  for (long start = 0; start < L; start += 1024L * 16) {
    const long p0 = start + (long)tid * 16;
    __align__(16) uint8_t b[16];
    unsigned bits = 0;
    if (p0 < L) {
      const int nv = (int)min(16L, L - p0);
      if (vec && nv == 16) {
        *(uint4*)b = *(const uint4*)(mp + p0);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) b[j] = j < nv ? mp[p0 + j] : (uint8_t)0;
      }
      uint8_t prev = p0 > 0 ? mp[p0 - 1] : b[0];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < nv && b[j] != prev) bits |= 1u << j;
        prev = b[j];
      }
    }
    const int n = __popc(bits);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;                             // exclusive prefix of the warp totals
    }
    __syncthreads();
    int off = base_sh + warp_tot[wid] + (incl - n);
    while (bits) {
      const int j = __ffs(bits) - 1;
      bits &= bits - 1;
      if (off < cap) out[off] = (int32_t)(p0 + j);
      ++off;
    }
    __syncthreads();
    if (tid == 1023) base_sh = off;                        // the last thread's end offset = running total
    __syncthreads();
  }
  if (tid == 0) cnt[k] = base_sh;
}

__global__ void stability_counts_kernel(const float* __restrict__ x, int32_t* __restrict__ counts, long P, float delta) {
  MS2_PDL_WAIT();
  const int nidx = blockIdx.y;
  const float* xp = x + (long)nidx * P;
  int hi = 0, lo = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long)gridDim.x * blockDim.x) {
    float v = xp[i];
    hi += v > delta;
    lo += v > -delta;
  }
  hi = __reduce_add_sync(0xffffffffu, hi);
  lo = __reduce_add_sync(0xffffffffu, lo);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counts + nidx * 2, hi);
    atomicAdd(counts + nidx * 2 + 1, lo);
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int ms2_axpby(const float* x, float a, const float* z, float b, float c, void* y, int y_dt, long n,
                         long zn, void* stream) {
  MS2_CHECK_ARG(x && y && n >= 0, "axpby: bad args");
  MS2_CHECK_ARG(!z || (zn > 0 && n % zn == 0), "axpby: z length must divide n");
  if (!n) return MS2_OK;
  const bool vec = n % 4 == 0 && n < (1L << 31) && (!z || zn % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
                   ((uintptr_t)y % 16 == 0) && (!z || (uintptr_t)z % 16 == 0);
  if (vec) {
    MS2_DISPATCH_DTYPE(y_dt, TO, (ms2_launch(axpby_vec4_kernel<TO>, grid_for(n / 4), 256, 0, ST, 
                                     x, a, z, b, c, (TO*)y, (unsigned)(n / 4), (unsigned)(z ? zn / 4 : 1))));
  } else {
    MS2_DISPATCH_DTYPE(y_dt, TO, (ms2_launch(axpby_kernel<TO>, grid_for(n), 256, 0, ST, x, a, z, b, c, (TO*)y, n, z ? zn : 1)));
  }
  MS2_CHECK_LAUNCH("axpby");
  return MS2_OK;
}
extern "C" int ms2_gate_rows(const float* x, const float* gate, float fill, float* y, int B, long P, void* stream) {
  MS2_CHECK_ARG(x && gate && y, "gate_rows: null");
  long n = (long)B * P;
  if (!n) return MS2_OK;
  ms2_launch(gate_rows_kernel, grid_for(n), 256, 0, ST, x, gate, fill, y, n, P);
  MS2_CHECK_LAUNCH("gate_rows");
  return MS2_OK;
}
extern "C" int ms2_select_plane(const float* x, const int32_t* idx, float* y, int B, int M, long P, void* stream) {
  MS2_CHECK_ARG(x && idx && y && M > 0, "select_plane: bad args");
  long n = (long)B * P;
  if (!n) return MS2_OK;
  ms2_launch(select_plane_kernel, grid_for(n), 256, 0, ST, x, idx, y, B, M, P);
  MS2_CHECK_LAUNCH("select_plane");
  return MS2_OK;
}
extern "C" int ms2_add_rowvec(const float* x, const float* v, float s, float* y, long M, int C, void* stream) {
  MS2_CHECK_ARG(x && v && y && M >= 0 && C > 0, "add_rowvec: bad args");
  if (!M) return MS2_OK;
  if (C % 4 == 0 && M * C < (1L << 31) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)v % 16 == 0))
    ms2_launch(axpby_vec4_kernel<float>, grid_for(M * C / 4), 256, 0, ST, x, 1.f, v, s, 0.f, y, (unsigned)(M * C / 4), (unsigned)(C / 4));
  else
    ms2_launch(add_rowvec_kernel, grid_for(M * C), 256, 0, ST, x, v, s, y, M * C, C);
  MS2_CHECK_LAUNCH("add_rowvec");
  return MS2_OK;
}
extern "C" int ms2_cast(const void* x, int x_dt, void* y, int y_dt, long n, void* stream) {
  MS2_CHECK_ARG(x && y && n >= 0, "cast: bad args");
  if (!n) return MS2_OK;
  int g = grid_for(n);
  const bool vec = n % 4 == 0 && n < (1L << 31) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  if (x_dt == MS2_F32 && y_dt == MS2_BF16 && vec)
    ms2_launch(cast_vec4_kernel<bf16>, grid_for(n / 4), 256, 0, ST, (const float*)x, (bf16*)y, (unsigned)(n / 4));
  else if (x_dt == MS2_F32 && y_dt == MS2_BF16) ms2_launch(cast_kernel<float, bf16>, g, 256, 0, ST, (const float*)x, (bf16*)y, n);
  else if (x_dt == MS2_BF16 && y_dt == MS2_F32) ms2_launch(cast_kernel<bf16, float>, g, 256, 0, ST, (const bf16*)x, (float*)y, n);
  else if (x_dt == MS2_F32 && y_dt == MS2_F32) ms2_launch(cast_kernel<float, float>, g, 256, 0, ST, (const float*)x, (float*)y, n);
  else if (x_dt == MS2_BF16 && y_dt == MS2_BF16) ms2_launch(cast_kernel<bf16, bf16>, g, 256, 0, ST, (const bf16*)x, (bf16*)y, n);
  else { ms2_set_error("cast: bad dtype"); return MS2_ERR_ARG; }
  MS2_CHECK_LAUNCH("cast");
  return MS2_OK;
}
extern "C" int ms2_activation(const float* x, float* y, long n, int act, void* stream) {
  MS2_CHECK_ARG(x && y && n >= 0, "activation: bad args");
  if (!n) return MS2_OK;
  ms2_launch(act_kernel, grid_for(n), 256, 0, ST, x, y, n, act);
  MS2_CHECK_LAUNCH("activation");
  return MS2_OK;
}
extern "C" int ms2_upsample2x_add(float* fine, const float* coarse, int B, int H, int W, int C, void* stream) {
  MS2_CHECK_ARG(fine && coarse && (H % 2 == 0) && (W % 2 == 0), "upsample2x_add: bad args");
  long n = (long)B * H * W * C;
  if (!n) return MS2_OK;
  ms2_launch(upsample2x_add_kernel, grid_for(n), 256, 0, ST, fine, coarse, B, H, W, C);
  MS2_CHECK_LAUNCH("upsample2x_add");
  return MS2_OK;
}
static int transpose_launch(const float* x, float* y, int B, long R, int Cc, cudaStream_t st) {
  if (!B || !R || !Cc) return MS2_OK;
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), B), block(32, 8);
  ms2_launch(transpose_kernel, grid, block, 0, st, x, y, R, Cc);
  MS2_CHECK_LAUNCH("transpose");
  return MS2_OK;
}
extern "C" int ms2_nhwc_to_nchw(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  MS2_CHECK_ARG(x && y, "nhwc_to_nchw: null");
  return transpose_launch(x, y, B, (long)H * W, C, ST);
}
extern "C" int ms2_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  MS2_CHECK_ARG(x && y, "nchw_to_nhwc: null");
  // input viewed as [C, HW] -> output [HW, C]
  if (!B) return MS2_OK;
  dim3 grid(ceil_div((long)H * W, 32), ceil_div(C, 32), B), block(32, 8);
  ms2_launch(transpose_kernel, grid, block, 0, ST, x, y, (long)C, H * W);
  MS2_CHECK_LAUNCH("transpose");
  return MS2_OK;
}
extern "C" int ms2_maxpool2x2(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  MS2_CHECK_ARG(x && y && (H % 2 == 0) && (W % 2 == 0), "maxpool2x2: bad args");
  long n = (long)B * (H / 2) * (W / 2) * C;
  if (!n) return MS2_OK;
  if (C % 4 == 0 && (long)B * H * W * C < (1L << 31) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0))
    ms2_launch(maxpool2x2_vec4_kernel, grid_for(n / 4), 256, 0, ST, x, y, (unsigned)(n / 4), (unsigned)H, (unsigned)W, (unsigned)(C / 4));
  else
    ms2_launch(maxpool2x2_kernel, grid_for(n), 256, 0, ST, x, y, B, H, W, C);
  MS2_CHECK_LAUNCH("maxpool2x2");
  return MS2_OK;
}
extern "C" int ms2_pixel_shuffle_add(const float* g, const float* bias, const float* skip, float* out, int B, int H,
                                     int W, int C, int act, void* stream) {
  MS2_CHECK_ARG(g && out, "pixel_shuffle_add: null");
  long n = (long)B * 4 * H * W * C;
  if (!n) return MS2_OK;
  if (C % 4 == 0 && n < (1L << 31) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
      (!bias || (uintptr_t)bias % 16 == 0) && (!skip || (uintptr_t)skip % 16 == 0))
    ms2_launch(pixel_shuffle_add_vec4_kernel, grid_for(n / 4), 256, 0, ST, g, bias, skip, out, (unsigned)(n / 4), (unsigned)H,
                                                                   (unsigned)W, (unsigned)(C / 4), act);
  else
    ms2_launch(pixel_shuffle_add_kernel, grid_for(n), 256, 0, ST, g, bias, skip, out, B, H, W, C, act);
  MS2_CHECK_LAUNCH("pixel_shuffle_add");
  return MS2_OK;
}
extern "C" int ms2_hyper_mask(const float* up, const float* hyper, float* masks, int B, int P, int C, int Mk,
                              void* stream) {
  MS2_CHECK_ARG(up && hyper && masks && C <= 64 && Mk <= 8, "hyper_mask: bad args (C<=64, Mk<=8)");
  if (!B || !P) return MS2_OK;
  if (C == 32 && Mk == 4 && ((uintptr_t)up % 16 == 0) && ((uintptr_t)hyper % 16 == 0)) {
    dim3 grid1((P + 127) / 128, B);
    ms2_launch(hyper_mask_c32m4_kernel, grid1, 128, 0, ST, up, hyper, masks, P);
  } else {
    dim3 grid(148 * 4, B);
    ms2_launch(hyper_mask_kernel, grid, 256, 0, ST, up, hyper, masks, B, P, C, Mk);
  }
  MS2_CHECK_LAUNCH("hyper_mask");
  return MS2_OK;
}
extern "C" int ms2_fourier_pe(const float* coords, const float* gauss, float* out, int n, int F, void* stream) {
  MS2_CHECK_ARG(coords && gauss && out, "fourier_pe: null");
  if (!n) return MS2_OK;
  ms2_launch(fourier_pe_kernel, grid_for((long)n * F), 256, 0, ST, coords, gauss, out, n, F);
  MS2_CHECK_LAUNCH("fourier_pe");
  return MS2_OK;
}
extern "C" int ms2_point_embed(const float* coords, const int* labels, const float* gauss, const float* table, float* out,
                               int B, int N, int pad, int F, int image_w, int image_h, const float* prefix, int n_prefix,
                               void* stream) {
  MS2_CHECK_ARG(coords && labels && gauss && table && out && N >= 0 && F > 0 && image_w > 0 && image_h > 0 && n_prefix >= 0,
                "point_embed: bad args");
  const long total = (long)B * (N + (pad ? 1 : 0) + (prefix ? n_prefix : 0)) * F;
  if (!total) return MS2_OK;
  ms2_launch(point_embed_kernel, grid_for(total), 256, 0, ST, coords, labels, gauss, table, out, B, N, pad, F, 1.f / image_w,
                                                      1.f / image_h, prefix, prefix ? n_prefix : 0);
  MS2_CHECK_LAUNCH("point_embed");
  return MS2_OK;
}
extern "C" int ms2_normalize_image(const void* x, int in_layout, void* out, int out_dt, int B, int H, int W, void* stream) {
  MS2_CHECK_ARG(x && out && in_layout >= 0 && in_layout <= 2, "normalize_image: bad args (layout 0 fp32 NCHW, 1 u8 NHWC, 2 u8 NCHW)");
  long n = (long)B * 3 * H * ((W + 3) / 4);
  if (!n) return MS2_OK;
  MS2_DISPATCH_DTYPE(out_dt, T, (ms2_launch(normalize_image_kernel<T>, grid_for(n), 256, 0, ST, x, in_layout, (T*)out, B, H, W)));
  MS2_CHECK_LAUNCH("normalize_image");
  return MS2_OK;
}
extern "C" int ms2_mask_stability_counts(const float* x, int32_t* counts, int N, long P, float delta, void* stream) {
  MS2_CHECK_ARG(x && counts, "mask_stability_counts: null");
  if (!N) return MS2_OK;
  MS2_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * N, ST), "stability memset");
  dim3 grid(64, N);
  ms2_launch(stability_counts_kernel, grid, 256, 0, ST, x, counts, P, delta);
  MS2_CHECK_LAUNCH("stability_counts");
  return MS2_OK;
}
extern "C" int ms2_seg_counts(const float* pred, const float* gt, const float* thr_host, int T, int32_t* counts, int N,
                              long P, void* stream) {
  MS2_CHECK_ARG(thr_host && T >= 1 && T <= 8 && N >= 0 && P >= 0, "seg_counts: bad args (1..8 thresholds)");
  if (!N) return MS2_OK;
  MS2_CHECK_ARG(counts, "seg_counts: counts is null");
  MS2_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * T * N, ST), "seg_counts memset");
  if (!P) return MS2_OK;
  MS2_CHECK_ARG(pred && gt, "seg_counts: null planes");
  MS2_CHECK_ARG(N <= 65535 && P < (1L << 31), "seg_counts: at most 65535 planes of < 2^31 pixels per call");
  SegThr thr;
  for (int t = 0; t < 8; ++t) thr.v[t] = t < T ? thr_host[t] : INFINITY;
  const bool vec = (P % 4 == 0) && (((uintptr_t)pred | (uintptr_t)gt) % 16 == 0);
  switch (T) {
    case 1: seg_counts_launch<1>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 2: seg_counts_launch<2>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 3: seg_counts_launch<3>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 4: seg_counts_launch<4>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 5: seg_counts_launch<5>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 6: seg_counts_launch<6>(vec, ST, pred, gt, thr, counts, N, P); break;
    case 7: seg_counts_launch<7>(vec, ST, pred, gt, thr, counts, N, P); break;
    default: seg_counts_launch<8>(vec, ST, pred, gt, thr, counts, N, P); break;
  }
  MS2_CHECK_LAUNCH("seg_counts");
  return MS2_OK;
}
extern "C" int ms2_bce_logits_sum(const float* pred, const float* gt, float pos_weight, double* sums, int N, long P,
                                  void* stream) {
  MS2_CHECK_ARG(N >= 0 && P >= 0, "bce_logits_sum: bad sizes");
  if (!N) return MS2_OK;
  MS2_CHECK_ARG(sums, "bce_logits_sum: sums is null");
  MS2_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * N, ST), "bce_logits_sum memset");
  if (!P) return MS2_OK;
  MS2_CHECK_ARG(pred && gt && N <= 65535, "bce_logits_sum: null planes or more than 65535 planes");
  const bool vec = (P % 4 == 0) && (((uintptr_t)pred | (uintptr_t)gt) % 16 == 0);
  long per_plane = (148L * 6) / N, by_size = (P + 8191) / 8192;      // one resident wave (6 CTAs of 256 threads per SM)
  if (per_plane > by_size) per_plane = by_size;
  dim3 grid((unsigned)(per_plane < 1 ? 1 : per_plane), N);
  if (vec) ms2_launch(bce_logits_sum_kernel<true>, grid, 256, 0, ST, pred, gt, pos_weight - 1.f, sums, P);
  else ms2_launch(bce_logits_sum_kernel<false>, grid, 256, 0, ST, pred, gt, pos_weight - 1.f, sums, P);
  MS2_CHECK_LAUNCH("bce_logits_sum");
  return MS2_OK;
}
extern "C" int ms2_mask_stats(const float* x, int32_t* stats, int N, int H, int W, float thr, float off, void* stream) {
  MS2_CHECK_ARG(N >= 0 && H >= 0 && W >= 0, "mask_stats: bad sizes");
  if (!N) return MS2_OK;
  MS2_CHECK_ARG(stats && N <= 65535, "mask_stats: stats is null or more than 65535 planes");
  ms2_launch(mask_stats_init_kernel, (N + 255) / 256, 256, 0, ST, stats, N, H, W);
  MS2_CHECK_LAUNCH("mask_stats_init");
  if (!H || !W) return MS2_OK;
  MS2_CHECK_ARG(x, "mask_stats: x is null");
  long per_plane = (148L * 8) / N;                        // one resident wave of 256-thread CTAs, at most one per row
  if (per_plane > H) per_plane = H;
  dim3 grid((unsigned)(per_plane < 1 ? 1 : per_plane), N);
  ms2_launch(mask_stats_kernel, grid, 256, 0, ST, x, stats, H, W, thr, off, W % 4 == 0 && (uintptr_t)x % 16 == 0);
  MS2_CHECK_LAUNCH("mask_stats");
  return MS2_OK;
}
extern "C" int ms2_mask_binarize_t(const float* x, const int32_t* sel, uint8_t* out, int K, int H, int W, float thr, int OH,
                                   int OW, int x0, int y0, void* stream) {
  MS2_CHECK_ARG(K >= 0 && H >= 0 && W >= 0 && x0 >= 0 && y0 >= 0 && x0 + W <= OW && y0 + H <= OH,
                "mask_binarize_t: the crop does not fit the output");
  if (!K || !OH || !OW) return MS2_OK;
  MS2_CHECK_ARG(out && K <= 65535, "mask_binarize_t: out is null or more than 65535 planes");
  if (H != OH || W != OW) MS2_CUDA(cudaMemsetAsync(out, 0, (size_t)K * OH * OW, ST), "mask_binarize_t memset");
  if (!H || !W) return MS2_OK;
  MS2_CHECK_ARG(x && sel, "mask_binarize_t: null input");
  dim3 grid((W + 31) / 32, (H + 31) / 32, K), block(32, 8);
  ms2_launch(mask_binarize_t_kernel, grid, block, 0, ST, x, sel, out, H, W, thr, OH, OW, x0, y0);
  MS2_CHECK_LAUNCH("mask_binarize_t");
  return MS2_OK;
}
extern "C" int ms2_rle_transitions(const uint8_t* m, int32_t* pos, int32_t* cnt, int K, long L, int cap, void* stream) {
  MS2_CHECK_ARG(K >= 0 && L >= 0 && L < (1L << 31) && cap >= 0, "rle_transitions: bad sizes");
  if (!K) return MS2_OK;
  MS2_CHECK_ARG(cnt && (pos || !cap) && (m || !L), "rle_transitions: null pointer");
  const bool vec = (L % 16 == 0) && ((uintptr_t)m % 16 == 0);
  ms2_launch(rle_transitions_kernel, K, 1024, 0, ST, m, L, pos, cnt, cap, vec);
  MS2_CHECK_LAUNCH("rle_transitions");
  return MS2_OK;
}
