// Per-slice glue of the tracking step as kernels (the reference does these with a dozen tiny torch ops each; on a
// latency-bound path every one of them is a dependent launch):
//   ms2_bank_rows           - sam2_base.py:566-637: the recent memories + object-pointer tokens of the current frame,
//                             concatenated, position-coded and cast in ONE pass into the memory bank's staging rows
//   ms2_argmax_select_rows  - sam2_base.py:383-395: best multimask by predicted IoU and its output token
//   ms2_obj_ptr_mix         - sam2_base.py:397-408: object pointer gated by the object score
//   ms2_stability_select    - mask_decoder.py:269-317: dynamic multimask via stability, index + IoU, no host sync
#include "common.cuh"

namespace {

constexpr int BR_MAX = 80;
struct BankRowsP {
  const float* src[BR_MAX];   // [B, rows, W] fp32 contiguous per source
  const float* pos[BR_MAX];   // [rows, W] (pos_bs == 0) or [B, rows, W] fp32, or null
  int rows[BR_MAX];
  int row0[BR_MAX];           // first destination row of the source
  long pos_bs[BR_MAX];
  int n, W, B;
  long k_bs, m_bs;            // destination batch strides (elements)
};

template <typename T>
__global__ void __launch_bounds__(256) bank_rows_kernel(const __grid_constant__ BankRowsP p, T* __restrict__ k_in,
                                                        T* __restrict__ m_out) {
  MS2_PDL_WAIT();
  const int s = blockIdx.y, b = blockIdx.z;
  const int rows = p.rows[s], W = p.W, W4 = W >> 2;
  const long n4 = (long)rows * W4;
  const float* src = p.src[s] + (long)b * rows * W;
  const float* pos = p.pos[s] ? p.pos[s] + (long)b * p.pos_bs[s] : nullptr;
  T* kd = k_in ? k_in + (long)b * p.k_bs + (long)p.row0[s] * W : nullptr;
  T* md = m_out ? m_out + (long)b * p.m_bs + (long)p.row0[s] * W : nullptr;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = __ldg((const float4*)src + i);
    float4 k = v;
    if (pos) {
      const float4 q = __ldg((const float4*)pos + i);
      k.x += q.x; k.y += q.y; k.z += q.z; k.w += q.w;
    }
    if (sizeof(T) == 4) {
      if (kd) ((float4*)kd)[i] = k;
      if (md) ((float4*)md)[i] = v;
    } else {
      if (kd) {
        __nv_bfloat162 a = __floats2bfloat162_rn(k.x, k.y), c = __floats2bfloat162_rn(k.z, k.w);
        ((uint2*)kd)[i] = make_uint2(*(uint32_t*)&a, *(uint32_t*)&c);
      }
      if (md) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), c = __floats2bfloat162_rn(v.z, v.w);
        ((uint2*)md)[i] = make_uint2(*(uint32_t*)&a, *(uint32_t*)&c);
      }
    }
  }
}

__global__ void argmax_select_rows_kernel(const float* __restrict__ scores, long srs, int M, const float* __restrict__ rows,
                                          long rows_bs, long rows_rs, int C, int32_t* __restrict__ idx_out,
                                          float* __restrict__ out) {
  MS2_PDL_WAIT();
  const int b = blockIdx.x;
  int best = 0;
  float bv = scores[(long)b * srs];
  for (int m = 1; m < M; ++m) {
    const float v = scores[(long)b * srs + m];
    if (v > bv) { bv = v; best = m; }                      // first maximum, as torch.argmax
  }
  if (threadIdx.x == 0 && idx_out) idx_out[b] = best;
  if (rows && out)
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[(long)b * C + c] = rows[(long)b * rows_bs + (long)best * rows_rs + c];
}

__global__ void obj_ptr_mix_kernel(const float* __restrict__ ptr, const float* __restrict__ logits,
                                   const float* __restrict__ no_obj, float* __restrict__ out, int C, int soft, int fixed) {
  MS2_PDL_WAIT();
  const int b = blockIdx.x;
  const float l = logits[b];
  const float lam = soft ? 1.f / (1.f + __expf(-l)) : (l > 0.f ? 1.f : 0.f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = ptr[(long)b * C + c];
    if (fixed) v = lam * v;
    out[(long)b * C + c] = v + (1.f - lam) * no_obj[c];
  }
}

__global__ void stability_select_kernel(const int32_t* __restrict__ counts, const float* __restrict__ ious, int M,
                                        float thresh, int32_t* __restrict__ idx_out, float* __restrict__ iou_out, int B) {
  MS2_PDL_WAIT();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float ai = (float)counts[2 * b], au = (float)counts[2 * b + 1];
  const float stab = au > 0.f ? ai / au : 1.f;
  int best = 1;
  float bv = ious[(long)b * M + 1];
  for (int m = 2; m < M; ++m) {
    const float v = ious[(long)b * M + m];
    if (v > bv) { bv = v; best = m; }
  }
  const int idx = stab >= thresh ? 0 : best;
  idx_out[b] = idx;
  iou_out[b] = ious[(long)b * M + idx];
}

}  // namespace

extern "C" int ms2_bank_rows(const void* const* h_src, const void* const* h_pos, const long* h_pos_bs, const int* h_rows,
                             int n, int W, int B, void* k_in, long k_bs, void* m_out, long m_bs, int out_dt, void* stream) {
  MS2_CHECK_ARG(n >= 0 && n <= BR_MAX, "bank_rows: at most %d sources per call (got %d)", BR_MAX, n);
  if (!n || !B) return MS2_OK;
  MS2_CHECK_ARG(h_src && h_rows && W > 0 && W % 4 == 0 && (k_in || m_out), "bank_rows: bad args (W must be a multiple of 4)");
  BankRowsP p;
  p.n = n; p.W = W; p.B = B; p.k_bs = k_bs; p.m_bs = m_bs;
  int row = 0, maxrows = 0;
  for (int i = 0; i < n; ++i) {
    MS2_CHECK_ARG(h_src[i] && h_rows[i] >= 0 && (uintptr_t)h_src[i] % 16 == 0, "bank_rows: source %d is null or misaligned", i);
    p.src[i] = (const float*)h_src[i];
    p.pos[i] = h_pos ? (const float*)h_pos[i] : nullptr;
    MS2_CHECK_ARG(!p.pos[i] || (uintptr_t)p.pos[i] % 16 == 0, "bank_rows: position table %d is misaligned", i);
    p.pos_bs[i] = h_pos_bs ? h_pos_bs[i] : 0;
    p.rows[i] = h_rows[i];
    p.row0[i] = row;
    row += h_rows[i];
    if (h_rows[i] > maxrows) maxrows = h_rows[i];
  }
  MS2_CHECK_ARG((!k_in || (uintptr_t)k_in % 16 == 0) && (!m_out || (uintptr_t)m_out % 16 == 0), "bank_rows: misaligned destination");
  long bx = ((long)maxrows * (W / 4) + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, n, B);
  if (out_dt == MS2_BF16) ms2_launch(bank_rows_kernel<bf16>, grid, 256, 0, (cudaStream_t)stream, p, (bf16*)k_in, (bf16*)m_out);
  else if (out_dt == MS2_F32) ms2_launch(bank_rows_kernel<float>, grid, 256, 0, (cudaStream_t)stream, p, (float*)k_in, (float*)m_out);
  else MS2_CHECK_ARG(false, "bank_rows: bad dtype");
  MS2_CHECK_LAUNCH("bank_rows");
  return MS2_OK;
}

extern "C" int ms2_argmax_select_rows(const float* scores, long scores_rs, int B, int M, const float* rows, long rows_bs, long rows_rs,
                                      int C, int32_t* idx_out, float* out, void* stream) {
  MS2_CHECK_ARG(scores && M >= 1 && (idx_out || out), "argmax_select_rows: bad args");
  if (!B) return MS2_OK;
  ms2_launch(argmax_select_rows_kernel, B, 128, 0, (cudaStream_t)stream, scores, scores_rs, M, rows, rows_bs, rows_rs, C, idx_out, out);
  MS2_CHECK_LAUNCH("argmax_select_rows");
  return MS2_OK;
}

extern "C" int ms2_obj_ptr_mix(const float* ptr, const float* logits, const float* no_obj, float* out, int B, int C,
                               int soft, int fixed, void* stream) {
  MS2_CHECK_ARG(ptr && logits && no_obj && out, "obj_ptr_mix: null pointer");
  if (!B) return MS2_OK;
  ms2_launch(obj_ptr_mix_kernel, B, 128, 0, (cudaStream_t)stream, ptr, logits, no_obj, out, C, soft, fixed);
  MS2_CHECK_LAUNCH("obj_ptr_mix");
  return MS2_OK;
}

extern "C" int ms2_stability_select(const int32_t* counts, const float* ious, int B, int M, float thresh, int32_t* idx_out,
                                    float* iou_out, void* stream) {
  MS2_CHECK_ARG(counts && ious && idx_out && iou_out && M >= 2, "stability_select: bad args");
  if (!B) return MS2_OK;
  ms2_launch(stability_select_kernel, (B + 127) / 128, 128, 0, (cudaStream_t)stream, counts, ious, M, thresh, idx_out, iou_out, B);
  MS2_CHECK_LAUNCH("stability_select");
  return MS2_OK;
}

// ------------------------------------------------------------------ several device-to-device copies in ONE launch
// The graph runner moves the inputs of a captured sub-pipeline into its static buffers and its results out of the graph's
// pool on every replay: 6-13 copies of 4 B..8 MB per slice, each 2-4 us of stream time as a separate memcpy node.
namespace {
constexpr int MC_MAX = 16;
struct MultiCopyP {
  const uint8_t* src[MC_MAX];
  uint8_t* dst[MC_MAX];
  long bytes[MC_MAX];
};
__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ MultiCopyP p) {
  MS2_PDL_WAIT();
  const int s = blockIdx.y;
  const uint8_t* src = p.src[s];
  uint8_t* dst = p.dst[s];
  const long n = p.bytes[s];
  const long stride = (long)gridDim.x * blockDim.x, t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const long n16 = n >> 4;
    for (long i = t; i < n16; i += stride) ((uint4*)dst)[i] = __ldg((const uint4*)src + i);
    for (long i = (n16 << 4) + t; i < n; i += stride) dst[i] = src[i];
  } else {
    for (long i = t; i < n; i += stride) dst[i] = src[i];
  }
}
}  // namespace

extern "C" int ms2_multi_copy(const void* const* h_src, void* const* h_dst, const long* h_bytes, int n, void* stream) {
  MS2_CHECK_ARG(n >= 0 && (n == 0 || (h_src && h_dst && h_bytes)), "multi_copy: null tables");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n; i0 += MC_MAX) {
    MultiCopyP p;
    int m = 0;
    long big = 0;
    for (int i = i0; i < n && m < MC_MAX; ++i) {
      MS2_CHECK_ARG(h_bytes[i] >= 0, "multi_copy: negative size");
      if (!h_bytes[i]) continue;
      MS2_CHECK_ARG(h_src[i] && h_dst[i], "multi_copy: null pointer in item %d", i);
      p.src[m] = (const uint8_t*)h_src[i];
      p.dst[m] = (uint8_t*)h_dst[i];
      p.bytes[m] = h_bytes[i];
      if (h_bytes[i] > big) big = h_bytes[i];
      ++m;
    }
    if (!m) continue;
    for (int i = m; i < MC_MAX; ++i) { p.src[i] = nullptr; p.dst[i] = nullptr; p.bytes[i] = 0; }
    long bx = (big / 16 + 255) / 256;                 // one 16-byte piece per thread for the largest item ...
    const long cap = (148L * 8 + m - 1) / m;          // ... capped so that the whole launch is ~8 CTAs per SM
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    ms2_launch(multi_copy_kernel, dim3((unsigned)bx, m), 256, 0, st, p);
    MS2_CHECK_LAUNCH("multi_copy");
  }
  return MS2_OK;
}
