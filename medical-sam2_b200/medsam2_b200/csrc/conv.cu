// Convolution-shaped kernels of the hot path that are NOT plain GEMMs:
//   patch_embed  — Hiera PatchEmbed 7x7/s4/p3, 3->96, NCHW image -> NHWC tokens (+pos-embed table)
//                  (backbones/utils.py:87-95, hieradet.py:279-284)
//   im2col       — tap gather for the small strided convs (mask down-sampler 3x3/s2, prompt-encoder
//                  2x2/s2, mask_downsample 4x4/s4) whose contraction then runs through ms2_gemm
//   dwconv7x7    — CXBlock depth-wise 7x7 (memory_encoder.py:84-90), NHWC, HBM/L2-bound
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ patch embed
constexpr int PE_T = 8;                      // 8x8 output pixels per CTA
constexpr int PE_IN = PE_T * 4 + 3;          // 35 input rows/cols
template <int COUT>
__global__ void __launch_bounds__(256)
patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ pos, float* __restrict__ out, int Hin, int Win, int Ho, int Wo) {
  MS2_PDL_WAIT();
  extern __shared__ float sm[];
  float* ws = sm;                                  // [147][COUT]
  float* patch = sm + 147 * COUT;                  // [3][35][36]
  constexpr int PS = PE_IN + 1;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * PE_T, ox0 = blockIdx.x * PE_T;
  for (int i = threadIdx.x; i < 147 * COUT; i += blockDim.x) {
    int co = i / 147, tap = i - co * 147;          // w is [COUT][3][7][7] = [COUT][147]
    ws[tap * COUT + co] = w[i];
  }
  const int iy0 = oy0 * 4 - 3, ix0 = ox0 * 4 - 3;
  for (int i = threadIdx.x; i < 3 * PE_IN * PE_IN; i += blockDim.x) {
    int ci = i / (PE_IN * PE_IN);
    int r = i - ci * PE_IN * PE_IN;
    int yy = r / PE_IN, xx = r - yy * PE_IN;
    int gy = iy0 + yy, gx = ix0 + xx;
    float v = 0.f;
    if (gy >= 0 && gy < Hin && gx >= 0 && gx < Win) v = img[(((long)b * 3 + ci) * Hin + gy) * Win + gx];
    patch[(ci * PE_IN + yy) * PS + xx] = v;
  }
  __syncthreads();
  constexpr int NCG = COUT / 4;                    // channel groups of 4
  const int t = threadIdx.x;
  if (t >= NCG * PE_T) return;
  const int cg = t % NCG, pg = t / NCG;            // pg = output row inside the tile
  float acc[4][PE_T];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < PE_T; ++j) acc[i][j] = 0.f;
  for (int ci = 0; ci < 3; ++ci)
    for (int ky = 0; ky < 7; ++ky) {
      const float* prow = patch + (ci * PE_IN + pg * 4 + ky) * PS;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + ((ci * 7 + ky) * 7 + kx) * COUT + cg * 4);
#pragma unroll
        for (int j = 0; j < PE_T; ++j) {
          const float x = prow[j * 4 + kx];
          acc[0][j] = fmaf(x, wv.x, acc[0][j]);
          acc[1][j] = fmaf(x, wv.y, acc[1][j]);
          acc[2][j] = fmaf(x, wv.z, acc[2][j]);
          acc[3][j] = fmaf(x, wv.w, acc[3][j]);
        }
      }
    }
  const int oy = oy0 + pg;
  if (oy >= Ho) return;
#pragma unroll
  for (int j = 0; j < PE_T; ++j) {
    const int ox = ox0 + j;
    if (ox >= Wo) continue;
    const long tok = ((long)oy * Wo + ox) * COUT + cg * 4;
    float4 o;
    o.x = acc[0][j] + bias[cg * 4 + 0];
    o.y = acc[1][j] + bias[cg * 4 + 1];
    o.z = acc[2][j] + bias[cg * 4 + 2];
    o.w = acc[3][j] + bias[cg * 4 + 3];
    if (pos) {
      const float4 pv = *reinterpret_cast<const float4*>(pos + tok);
      o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
    }
    *reinterpret_cast<float4*>(out + (long)b * Ho * Wo * COUT + tok) = o;
  }
}

// ------------------------------------------------------------------ im2col
template <typename T>
__global__ void im2col_kernel(const float* __restrict__ x, T* __restrict__ cols, int B, int H, int W, int Cin, int k,
                              int stride, int pad, int Ho, int Wo, int pre, float pre_scale, float pre_bias) {
  MS2_PDL_WAIT();
  const int KK = k * k * Cin;
  const long n = (long)B * Ho * Wo * KK;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int col = i % KK;
    long t = i / KK;
    int xo = t % Wo; t /= Wo;
    int yo = t % Ho;
    int b = t / Ho;
    int ci = col % Cin;
    int tap = col / Cin;
    int ky = tap / k, kx = tap - ky * k;
    int y = yo * stride - pad + ky, xx = xo * stride - pad + kx;
    float v = 0.f;
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
      v = x[(((long)b * H + y) * W + xx) * Cin + ci];
      if (pre == 1) v = 1.f / (1.f + expf(-v));
      else if (pre == 2) v = v > 0.f ? 1.f : 0.f;
      v = v * pre_scale + pre_bias;
    }
    cols[i] = from_f<T>(v);
  }
}

// vectorised im2col for Cin % 8 == 0: one thread = 8 consecutive channels of one tap (16-byte bf16 store)
__global__ void __launch_bounds__(256)
im2col_vec8_kernel(const float* __restrict__ x, bf16* __restrict__ cols, long total, int H, int W, int Cin, int k,
                   int stride, int pad, int Ho, int Wo) {
  MS2_PDL_WAIT();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = Cin >> 3;
  const int cc = (int)(t % c8);
  long r = t / c8;
  const int tap = (int)(r % (k * k));
  r /= (k * k);
  const int xo = (int)(r % Wo);
  r /= Wo;
  const int yo = (int)(r % Ho), b = (int)(r / Ho);
  const int ky = tap / k, kx = tap - ky * k;
  const int y = yo * stride - pad + ky, xx = xo * stride - pad + kx;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (y >= 0 && y < H && xx >= 0 && xx < W) {
    const float4* src = (const float4*)(x + (((long)b * H + y) * W + xx) * Cin + cc * 8);
    const float4 a = src[0], c = src[1];
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
    u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
  }
  *(uint4*)(cols + t * 8) = u;
}

// ------------------------------------------------------------------ fused conv3x3/s2/p1 + LayerNorm2d + GELU
// The first layers of the mask down-sampler (memory_encoder.py:38-58: 1->4->16->64 channels) are far too thin for
// a GEMM: one thread computes ALL output channels of one output pixel (weights broadcast from shared memory),
// applies the channel LayerNorm and GELU in registers and writes the NHWC row - replacing im2col + GEMM + LN.
// pre: 0 none, 1 sigmoid, 2 (x > 0); then x*pre_scale + pre_bias on in-range inputs (sam2_base.py:686-696).
template <int CIN, int COUT, typename TO>
__global__ void __launch_bounds__(128)
conv3x3s2_ln_gelu_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ gamma, const float* __restrict__ beta, TO* __restrict__ y, int B,
                         int H, int W, int Ho, int Wo, float eps, int pre, float pre_scale, float pre_bias) {
  MS2_PDL_WAIT();
  __shared__ __align__(16) float ws[9 * CIN * COUT];   // [tap][ci][co]
  __shared__ float sb[3 * COUT];
  // conv weight [Cout, Cin, 3, 3] -> [tap][ci][co]: read in SOURCE order with 16-byte loads that are all issued before
  // the first store (the element-wise gather paid one global round trip per pass: 72 passes for 16 -> 64 channels,
  // most of the kernel's 65 us), scatter into shared memory
  {
    constexpr int NW = 9 * CIN * COUT, NV = NW / 4, PASSES = (NV + 127) / 128;
    static_assert(NW % 4 == 0, "weight count must be a multiple of 4");
    float4 v[PASSES];
#pragma unroll
    for (int u = 0; u < PASSES; ++u) {
      const int i4 = threadIdx.x + u * 128;
      if (i4 < NV) v[u] = __ldg((const float4*)w + i4);
    }
#pragma unroll
    for (int u = 0; u < PASSES; ++u) {
      const int i4 = threadIdx.x + u * 128;
      if (i4 < NV) {
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i4 * 4 + k;
          const int tap = i % 9, ci = (i / 9) % CIN, co = i / (9 * CIN);
          ws[(tap * CIN + ci) * COUT + co] = e[k];
        }
      }
    }
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    sb[i] = bias[i];
    sb[COUT + i] = gamma[i];
    sb[2 * COUT + i] = beta[i];
  }
  __syncthreads();
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long)B * Ho * Wo) return;
  const int xo = (int)(p % Wo);
  const long r = p / Wo;
  const int yo = (int)(r % Ho), b = (int)(r / Ho);
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = yo * 2 - 1 + tap / 3, xx = xo * 2 - 1 + tap % 3;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* src = x + (((long)b * H + yy) * W + xx) * CIN;
    float in[CIN];
    if (CIN % 4 == 0) {                                  // a pixel's channels as 16-byte loads
#pragma unroll
      for (int c4 = 0; c4 < CIN / 4; ++c4) {
        const float4 t = *(const float4*)(src + c4 * 4);
        in[c4 * 4] = t.x; in[c4 * 4 + 1] = t.y; in[c4 * 4 + 2] = t.z; in[c4 * 4 + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) in[ci] = src[ci];
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float v = in[ci];
      if (pre == 1) v = 1.f / (1.f + expf(-v));
      else if (pre == 2) v = v > 0.f ? 1.f : 0.f;
      if (pre) v = v * pre_scale + pre_bias;
      const float4* wr = (const float4*)(ws + (tap * CIN + ci) * COUT);      // broadcast 16-byte reads
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 wv = wr[c4];
        acc[c4 * 4] = fmaf(v, wv.x, acc[c4 * 4]);
        acc[c4 * 4 + 1] = fmaf(v, wv.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(v, wv.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(v, wv.w, acc[c4 * 4 + 3]);
      }
    }
  }
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < COUT; ++c) mean += acc[c];
  mean *= (1.f / COUT);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    const float d = acc[c] - mean;
    var = fmaf(d, d, var);
  }
  const float rstd = rsqrtf(var * (1.f / COUT) + eps);
  TO* dst = y + p * COUT;
  if (sizeof(TO) == 4) {
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c4 * 4 + k;
        o[k] = gelu_erf((acc[c] - mean) * rstd * sb[COUT + c] + sb[2 * COUT + c]);
      }
      *(float4*)((float*)dst + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < COUT; ++c) dst[c] = from_f<TO>(gelu_erf((acc[c] - mean) * rstd * sb[COUT + c] + sb[2 * COUT + c]));
  }
}

// ------------------------------------------------------------------ depth-wise 7x7
// HBM-bound (2 x 4 B x pixels x C): an 8x8 output tile x 32 channels per CTA; the 14x14 input halo tile is staged
// once in shared memory ([row][col][channel]: channel = bank, conflict-free) and each thread slides a 7-tap window
// along one output row of one channel, so every input value is read from HBM once (+ halo) instead of 49 times.
constexpr int DW_T = 8, DW_C = 32, DW_IN = DW_T + 6;
__global__ void __launch_bounds__(DW_T * DW_C)
dwconv7x7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, int B, int H, int W, int C) {
  MS2_PDL_WAIT();
  __shared__ __align__(16) float tile[DW_IN][DW_IN][DW_C];
  __shared__ float ws[49][DW_C];
  const int tilesx = (W + DW_T - 1) / DW_T;
  const int tx0 = (blockIdx.x % tilesx) * DW_T, ty0 = (blockIdx.x / tilesx) * DW_T;
  const int c0 = blockIdx.y * DW_C, b = blockIdx.z;
  const int cl = threadIdx.x % DW_C, ty = threadIdx.x / DW_C;
  for (int i = threadIdx.x; i < 49 * DW_C; i += blockDim.x) {
    const int c = i % DW_C, tap = i / DW_C;
    ws[tap][c] = (c0 + c < C) ? w[(long)(c0 + c) * 49 + tap] : 0.f;
  }
  if (c0 + DW_C <= C && C % 4 == 0 && ((uintptr_t)x % 16 == 0)) {
    // halo tile as asynchronous 16-byte copies (a pixel's 32 channels are one 128-byte line): every piece of a
    // thread is in flight at once; out-of-image pixels are zero-filled with ordinary stores
    for (int i = threadIdx.x; i < DW_IN * DW_IN * (DW_C / 4); i += blockDim.x) {
      const int c4 = i % (DW_C / 4), col = (i / (DW_C / 4)) % DW_IN, row = i / ((DW_C / 4) * DW_IN);
      const int sy = ty0 + row - 3, sx = tx0 + col - 3;
      float* dst = &tile[row][col][c4 * 4];
      if (sy >= 0 && sy < H && sx >= 0 && sx < W)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                     "l"(x + (((long)b * H + sy) * W + sx) * C + c0 + c4 * 4) : "memory");
      else
        *(float4*)dst = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int i = threadIdx.x; i < DW_IN * DW_IN * DW_C; i += blockDim.x) {
      const int c = i % DW_C, col = (i / DW_C) % DW_IN, row = i / (DW_C * DW_IN);
      const int sy = ty0 + row - 3, sx = tx0 + col - 3;
      float v = 0.f;
      if (sy >= 0 && sy < H && sx >= 0 && sx < W && c0 + c < C) v = x[(((long)b * H + sy) * W + sx) * C + c0 + c];
      tile[row][col][c] = v;
    }
  }
  __syncthreads();
  const int c = c0 + cl;
  float acc[DW_T];
  const float bv = (bias && c < C) ? bias[c] : 0.f;
#pragma unroll
  for (int i = 0; i < DW_T; ++i) acc[i] = bv;
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
    float in[DW_IN];
#pragma unroll
    for (int i = 0; i < DW_IN; ++i) in[i] = tile[ty + ky][i][cl];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float wv = ws[ky * 7 + kx][cl];
#pragma unroll
      for (int i = 0; i < DW_T; ++i) acc[i] = fmaf(in[i + kx], wv, acc[i]);
    }
  }
  const int oy = ty0 + ty;
  if (c < C && oy < H) {
#pragma unroll
    for (int i = 0; i < DW_T; ++i)
      if (tx0 + i < W) y[(((long)b * H + oy) * W + tx0 + i) * C + c] = acc[i];
  }
}

// ------------------------------------------------------------------ patch-embed im2col (bf16 path)
// NCHW fp32 image -> bf16 rows [B*Ho*Wo, ldk] of the 7x7/s4/p3 taps in (ky, kx, c) order, zero beyond 147; the
// contraction itself then runs on the tensor-core GEMM (K = 147 padded to a 16-byte multiple).  One thread = one
// 16-byte chunk of a row: coalesced stores, gathers served from L1.
template <typename IN>
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const IN* __restrict__ img, bf16* __restrict__ cols, long total_chunks, int Hin, int Win, int Ho,
                    int Wo, int chunks) {
  MS2_PDL_WAIT();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_chunks) return;
  const int ch = (int)(t % chunks);
  const long row = t / chunks;
  const int ox = (int)(row % Wo);
  const long r2 = row / Wo;
  const int oy = (int)(r2 % Ho), b = (int)(r2 / Ho);
  const IN* base = img + (long)b * 3 * Hin * Win;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int idx = ch * 8 + e;
    float val = 0.f;
    if (idx < 147) {
      const int ky = idx / 21, r = idx - ky * 21, kx = r / 3, c = r - kx * 3;
      const int sy = oy * 4 + ky - 3, sx = ox * 4 + kx - 3;
      if (sy >= 0 && sy < Hin && sx >= 0 && sx < Win) val = to_f(base[((long)c * Hin + sy) * Win + sx]);
    }
    v[e] = val;
  }
  uint4 u;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
  *(uint4*)(cols + t * 8) = u;
}

// Tiled variant: one CTA = 64 consecutive output pixels of one output row.  The 7 input rows x 3 channels x 259 pixels
// they cover are staged in shared memory with coalesced loads (the per-thread gather above issues 8 scattered scalar loads
// per 16 bytes written and runs at a quarter of the HBM rate), then written out as whole 16-byte column chunks.
constexpr int PT_OX = 64, PT_W = PT_OX * 4 + 3;
template <typename IN>
__global__ void __launch_bounds__(256)
patch_im2col_tiled_kernel(const IN* __restrict__ img, bf16* __restrict__ cols, int Hin, int Win, int Ho, int Wo, int chunks) {
  MS2_PDL_WAIT();
  __shared__ float tile[3][7][PT_W + 1];
  const int ox0 = blockIdx.x * PT_OX, oy = blockIdx.y, b = blockIdx.z;
  const IN* base = img + (long)b * 3 * Hin * Win;
  const int x_base = ox0 * 4 - 3, y_base = oy * 4 - 3;
  for (int i = threadIdx.x; i < 3 * 7 * PT_W; i += 256) {
    const int x = i % PT_W, r = i / PT_W, ky = r % 7, c = r / 7;
    const int sy = y_base + ky, sx = x_base + x;
    tile[c][ky][x] = (sy >= 0 && sy < Hin && sx >= 0 && sx < Win) ? to_f(base[((long)c * Hin + sy) * Win + sx]) : 0.f;
  }
  __syncthreads();
  const long row0 = ((long)b * Ho + oy) * Wo + ox0;
  for (int i = threadIdx.x; i < PT_OX * chunks; i += 256) {
    const int ch = i % chunks, o = i / chunks;
    if (ox0 + o >= Wo) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = ch * 8 + e;
      float val = 0.f;
      if (idx < 147) {
        const int ky = idx / 21, r = idx - ky * 21, kx = r / 3, c = r - kx * 3;
        val = tile[c][ky][o * 4 + kx];
      }
      v[e] = val;
    }
    uint4 u;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
    *(uint4*)(cols + ((row0 + o) * chunks + ch) * 8) = u;
  }
}

}  // namespace

extern "C" int ms2_conv3x3s2_ln_gelu(const float* x, const float* w, const float* bias, const float* gamma,
                                     const float* beta, void* y, int y_dt, int B, int H, int W, int Cin, int Cout,
                                     float eps, int pre, float pre_scale, float pre_bias, void* stream) {
  MS2_CHECK_ARG(x && w && bias && gamma && beta && y, "conv3x3s2_ln_gelu: null pointer");
  MS2_CHECK_ARG((uintptr_t)w % 16 == 0, "conv3x3s2_ln_gelu: w must be 16-byte aligned");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long np = (long)B * Ho * Wo;
  if (!np) return MS2_OK;
  const int grid = ceil_div(np, 128);
  cudaStream_t st = (cudaStream_t)stream;
#define MS2_C3(CI, CO)                                                                                               \
  do {                                                                                                               \
    if (y_dt == MS2_F32)                                                                                             \
      ms2_launch(conv3x3s2_ln_gelu_kernel<CI, CO, float>, grid, 128, 0, st, x, w, bias, gamma, beta, (float*)y, B, H, W, Ho, Wo, \
                                                                    eps, pre, pre_scale, pre_bias);                  \
    else                                                                                                             \
      ms2_launch(conv3x3s2_ln_gelu_kernel<CI, CO, bf16>, grid, 128, 0, st, x, w, bias, gamma, beta, (bf16*)y, B, H, W, Ho, Wo,  \
                                                                   eps, pre, pre_scale, pre_bias);                   \
  } while (0)
  if (Cin == 1 && Cout == 4) MS2_C3(1, 4);
  else if (Cin == 4 && Cout == 16) MS2_C3(4, 16);
  else if (Cin == 16 && Cout == 64) MS2_C3(16, 64);
  else {
    ms2_set_error("conv3x3s2_ln_gelu: only (1->4), (4->16), (16->64) are instantiated (got %d->%d)", Cin, Cout);
    return MS2_ERR_UNSUPPORTED;
  }
#undef MS2_C3
  MS2_CHECK_LAUNCH("conv3x3s2_ln_gelu_kernel");
  return MS2_OK;
}

extern "C" int ms2_patch_im2col(const void* img, int img_dt, void* cols, int B, int Hin, int Win, int ldk, void* stream) {
  MS2_CHECK_ARG(img && cols, "patch_im2col: null pointer");
  MS2_CHECK_ARG(ldk >= 152 && ldk % 8 == 0 && ((uintptr_t)cols % 16 == 0), "patch_im2col: ldk must be a multiple of 8 >= 152");
  const int Ho = (Hin + 6 - 7) / 4 + 1, Wo = (Win + 6 - 7) / 4 + 1;
  const long total = (long)B * Ho * Wo * (ldk / 8);
  if (!total) return MS2_OK;
  if (Ho <= 65535 && B <= 65535) {
    dim3 grid(ceil_div(Wo, PT_OX), Ho, B);
    MS2_DISPATCH_DTYPE(img_dt, T, (ms2_launch(patch_im2col_tiled_kernel<T>, grid, 256, 0, (cudaStream_t)stream, (const T*)img,
                                              (bf16*)cols, Hin, Win, Ho, Wo, ldk / 8)));
    MS2_CHECK_LAUNCH("patch_im2col_tiled_kernel");
    return MS2_OK;
  }
  MS2_DISPATCH_DTYPE(img_dt, T, (ms2_launch(patch_im2col_kernel<T>, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, 
                                    (const T*)img, (bf16*)cols, total, Hin, Win, Ho, Wo, ldk / 8)));
  MS2_CHECK_LAUNCH("patch_im2col_kernel");
  return MS2_OK;
}

extern "C" int ms2_patch_embed(const float* img, const float* w, const float* bias, const float* pos, float* out,
                               int B, int Hin, int Win, int Cout, void* stream) {
  MS2_CHECK_ARG(img && w && bias && out, "patch_embed: null pointer");
  MS2_CHECK_ARG(Cout == 96, "patch_embed: only embed_dim 96 is instantiated (got %d)", Cout);
  const int Ho = (Hin + 6 - 7) / 4 + 1, Wo = (Win + 6 - 7) / 4 + 1;
  if (!B) return MS2_OK;
  size_t smem = sizeof(float) * (147 * 96 + 3 * PE_IN * (PE_IN + 1));
  MS2_CUDA(cudaFuncSetAttribute(patch_embed_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
           "patch_embed attr");
  dim3 grid(ceil_div(Wo, PE_T), ceil_div(Ho, PE_T), B);
  ms2_launch(patch_embed_kernel<96>, grid, 256, smem, (cudaStream_t)stream, img, w, bias, pos, out, Hin, Win, Ho, Wo);
  MS2_CHECK_LAUNCH("patch_embed");
  return MS2_OK;
}

extern "C" int ms2_im2col(const float* x, void* cols, int dt, int B, int H, int W, int Cin, int k, int stride, int pad,
                          int pre, float pre_scale, float pre_bias, void* stream) {
  MS2_CHECK_ARG(x && cols && k > 0 && stride > 0, "im2col: bad args");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  long n = (long)B * Ho * Wo * k * k * Cin;
  if (!n) return MS2_OK;
  if (dt == MS2_BF16 && pre == 0 && Cin % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)cols % 16 == 0)) {
    const long total = n / 8;
    ms2_launch(im2col_vec8_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, x, (bf16*)cols, total, H, W, Cin, k, stride,
                                                                               pad, Ho, Wo);
    MS2_CHECK_LAUNCH("im2col_vec8");
    return MS2_OK;
  }
  long blocks = (n + 255) / 256;
  int g = (int)(blocks > 148L * 32 ? 148L * 32 : blocks);
  MS2_DISPATCH_DTYPE(dt, T, (ms2_launch(im2col_kernel<T>, g, 256, 0, (cudaStream_t)stream, 
                                x, (T*)cols, B, H, W, Cin, k, stride, pad, Ho, Wo, pre, pre_scale, pre_bias)));
  MS2_CHECK_LAUNCH("im2col");
  return MS2_OK;
}

extern "C" int ms2_dwconv7x7(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int C,
                             void* stream) {
  MS2_CHECK_ARG(x && w && y, "dwconv7x7: null pointer");
  long npix = (long)B * H * W;
  if (!npix) return MS2_OK;
  dim3 grid(ceil_div(W, DW_T) * ceil_div(H, DW_T), ceil_div(C, DW_C), B);
  ms2_launch(dwconv7x7_kernel, grid, DW_T * DW_C, 0, (cudaStream_t)stream, x, w, bias, y, B, H, W, C);
  MS2_CHECK_LAUNCH("dwconv7x7");
  return MS2_OK;
}
