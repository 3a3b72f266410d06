// Axial rotary position encoding applied in place (position_encoding.py:167-216; call site
// transformer.py:301-315).  Adjacent channel pairs (2i,2i+1) rotate by the table angle of the token's
// position inside its 64x64 frame; memory keys restart at 0 every table_len rows (rope_k_repeat) and
// the trailing object-pointer rows are not rotated.  fp32 math, like the reference's complex fp32.
#include "common.cuh"

namespace {
template <typename T>
__global__ void rope_kernel(T* __restrict__ x, long batch_stride, long row_stride, int B, int rows, int n_rope_rows,
                            int D, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int table_len) {
  MS2_PDL_WAIT();
  const int half = D / 2;
  const long n = (long)B * n_rope_rows * half;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int pr = i % half;
    long t = i / half;
    int r = t % n_rope_rows;
    int b = t / n_rope_rows;
    int pos = r % table_len;
    float c = cos_t[(long)pos * half + pr], s = sin_t[(long)pos * half + pr];
    T* p = x + b * batch_stride + r * row_stride + 2 * pr;
    float a = to_f(p[0]), bb = to_f(p[1]);
    p[0] = from_f<T>(a * c - bb * s);
    p[1] = from_f<T>(a * s + bb * c);
  }
}
// bf16, four pairs (16 bytes) per thread, 32-bit index arithmetic, 16-byte table loads
__global__ void rope_bf16_vec_kernel(bf16* __restrict__ x, long batch_stride, long row_stride, unsigned n4,
                                     unsigned n_rope_rows, unsigned q4, const float* __restrict__ cos_t,
                                     const float* __restrict__ sin_t, unsigned table_len) {
  MS2_PDL_WAIT();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned p4 = i % q4;                       // group of 4 pairs inside the row
    unsigned t = i / q4;
    const unsigned r = t % n_rope_rows, b = t / n_rope_rows;
    const unsigned pos = r % table_len;
    const float4 c = *(const float4*)(cos_t + ((long)pos * q4 + p4) * 4);
    const float4 s = *(const float4*)(sin_t + ((long)pos * q4 + p4) * 4);
    uint4* ptr = (uint4*)(x + b * batch_stride + (long)r * row_stride + p4 * 8);
    uint4 v = *ptr;
    __nv_bfloat162* h = (__nv_bfloat162*)&v;
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 ab = __bfloat1622float2(h[k]);
      h[k] = __floats2bfloat162_rn(ab.x * cc[k] - ab.y * ss[k], ab.x * ss[k] + ab.y * cc[k]);
    }
    *ptr = v;
  }
}
}  // namespace

extern "C" int ms2_rope(void* x, int dt, long batch_stride, long row_stride, int B, int rows, int n_rope_rows, int D,
                        const float* cos_t, const float* sin_t, int table_len, void* stream) {
  MS2_CHECK_ARG(x && cos_t && sin_t && D % 2 == 0 && table_len > 0, "rope: bad args");
  MS2_CHECK_ARG(n_rope_rows <= rows, "rope: n_rope_rows > rows");
  long n = (long)B * n_rope_rows * (D / 2);
  if (n <= 0) return MS2_OK;
  if (dt == MS2_BF16 && D % 8 == 0 && row_stride % 8 == 0 && batch_stride % 8 == 0 && ((uintptr_t)x % 16 == 0) &&
      ((uintptr_t)cos_t % 16 == 0) && ((uintptr_t)sin_t % 16 == 0) && n < (1L << 31)) {
    const long n4 = n / 4;
    long b4 = (n4 + 255) / 256;
    if (b4 > 148L * 32) b4 = 148L * 32;
    ms2_launch(rope_bf16_vec_kernel, (int)b4, 256, 0, (cudaStream_t)stream, (bf16*)x, batch_stride, row_stride, (unsigned)n4,
                                                                   (unsigned)n_rope_rows, (unsigned)(D / 8), cos_t, sin_t,
                                                                   (unsigned)table_len);
    MS2_CHECK_LAUNCH("rope");
    return MS2_OK;
  }
  long blocks = (n + 255) / 256;
  int g = (int)(blocks > 148L * 32 ? 148L * 32 : blocks);
  MS2_DISPATCH_DTYPE(dt, T, (ms2_launch(rope_kernel<T>, g, 256, 0, (cudaStream_t)stream, 
                                (T*)x, batch_stride, row_stride, B, rows, n_rope_rows, D, cos_t, sin_t, table_len)));
  MS2_CHECK_LAUNCH("rope");
  return MS2_OK;
}
