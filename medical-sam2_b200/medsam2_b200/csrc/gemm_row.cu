// Row-complete tcgen05 GEMM tiles (128 rows x 256 columns = one whole row of the 256-wide memory-attention stream per
// thread) with the epilogues that otherwise cost the tracked-frame path a kernel each:
//   ms2_gemm_res_ln : x = residual + A W^T + b (fp32, the residual stream)  AND  t = LayerNorm(x) gamma + beta (bf16 / fp32)
//                     reference: memory_attention.py:58-99 (`tgt = tgt + dropout(tgt2); tgt2 = self.norm_k(tgt)`) and :166
//   ms2_gemm_rope   : q = RoPE(A W^T + b) (bf16) for the column tiles below `rope_cols`, plain bf16 for the others
//                     reference: transformer.py:288-318 (q_proj / k_proj followed by apply_rotary_enc)
// The memory-attention stack of a tracked frame is a chain of ~13 dependent kernels of 5-15 us per layer on 4096 rows.
// M = 4096 rows give 32 CTAs (x 3 column tiles for the fused q|k|v projection); two epilogue warps per TMEM lane quarter
// own 128 columns of a row each.  MEASURED: correct (tests/test_gpu_kernels.py) but slower in the stack than the chain of
// separate kernels (409 vs 421 slices/s, see csrc/memattn.cu: 32 CTAs are bound by one SM's TMA ingest), so the native
// driver only uses these tiles when MS2_MEMATTN_FUSED=1.
//   LN:   pass 1  v = acc + bias + residual -> x (global, fp32) and back into tensor memory, row sum
//         pass 2  centred second moment     pass 3  (v - mean) rstd gamma + beta -> t
//   RoPE: v = bf16(acc + bias) (the rounding autocast applies to the projection), adjacent pairs rotated in fp32 by the
//         table angle of the row's position, rounded to bf16 again (position_encoding.py:185-216)
#include "tc_common.cuh"

namespace {

constexpr int RBM = 128, RBN = 256, RBK = 64, RST = 4;
constexpr int R_A_BYTES = RBM * RBK * 2, R_W_BYTES = RBN * RBK * 2, R_STAGE = R_A_BYTES + R_W_BYTES;
constexpr int R_SMEM = RST * R_STAGE + 1024 + 256 + 2048;
constexpr int R_THREADS = 320;       // producer, MMA issuer, 8 epilogue warps (two per TMEM lane quarter, 128 columns each)

struct RowP {
  const float* bias;        // [N] or null
  const float* residual;    // LN: fp32 [M, ldr]
  long ldr;
  float* x_out;             // LN: fp32 [M, ldx] (may alias residual)
  long ldx;
  void* t_out;              // LN: bf16 / fp32 [M, ldt]; RoPE: bf16 [M, ldt]
  long ldt;
  int t_f32;
  const float* gamma;
  const float* beta;
  float eps;
  const float* cos_t;       // [rope_len, 128]
  const float* sin_t;
  int rope_len, rope_tiles, L;
  int M, N, K, num_kb;
};

template <int MODE>          // 0: bf16 output, RoPE on column tiles < rope_tiles;  1: residual + LayerNorm
__global__ void __launch_bounds__(R_THREADS, 1)
gemm_row_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const RowP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + RST * R_STAGE);
  uint64_t* empty = full + RST;
  uint64_t* acc_full = empty + RST;
  uint32_t* tmem_ptr = (uint32_t*)(acc_full + 1);
  float* red = (float*)(smem + RST * R_STAGE + 256);      // [2 passes][2 halves][128 rows] LayerNorm partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * RBM, n0 = blockIdx.y * RBN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    for (int s = 0; s < RST; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, RBN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % RST;
        tc::mbar_wait(&empty[s], ((uint32_t)(kb / RST) & 1u) ^ 1u);
        uint8_t* sa = smem + (size_t)s * R_STAGE;
        tc::mbar_arrive_expect_tx(&full[s], R_STAGE);
        tc::tma_load_2d(sa, &tmA, &full[s], kb * RBK, m0);
        tc::tma_load_2d(sa + R_A_BYTES, &tmW, &full[s], kb * RBK, n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(RBM, RBN, 0, 0);
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % RST;
      tc::mbar_wait(&full[s], (uint32_t)(kb / RST) & 1u);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t sa = tc::smem_u32(smem + (size_t)s * R_STAGE), sb = sa + R_A_BYTES;
#pragma unroll
        for (int k = 0; k < RBK / 16; ++k)
          tc::umma_bf16(tmem_base, tc::desc_kmajor_sw128(sa + k * 32), tc::desc_kmajor_sw128(sb + k * 32), idesc,
                        (kb | k) ? 1u : 0u);
        if (kb + RST < p.num_kb) tc::umma_commit(&empty[s]);
        if (kb == p.num_kb - 1) tc::umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warps 2..9; a thread owns 128 columns of one row =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = m0 + q * 32 + lane;                         // M is a multiple of 128: every row exists
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
    constexpr int NCH = 4;                                     // 32-column chunks per thread
    if (MODE == 1) {
      const float* res = p.residual + (long)r * p.ldr + half * 128;
      float* xo = p.x_out + (long)r * p.ldx + half * 128;
      const float* bias = p.bias ? p.bias + half * 128 : nullptr;
      // the residual row does not depend on the accumulator: chunk c+1 is fetched while chunk c is processed, chunk 0
      // before the accumulator is even waited for
      float4 rr[8], rn[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) rr[g] = *(const float4*)(res + g * 4);
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t acc[32];
        tc::tmem_ld32(taddr + c * 32, acc);
        if (c + 1 < NCH) {
#pragma unroll
          for (int g = 0; g < 8; ++g) rn[g] = *(const float4*)(res + (c + 1) * 32 + g * 4);
        }
        float4 bb[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) bb[g] = bias ? __ldg((const float4*)(bias + c * 32 + g * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        tc::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 v;
          v.x = __uint_as_float(acc[g * 4]) + bb[g].x + rr[g].x;
          v.y = __uint_as_float(acc[g * 4 + 1]) + bb[g].y + rr[g].y;
          v.z = __uint_as_float(acc[g * 4 + 2]) + bb[g].z + rr[g].z;
          v.w = __uint_as_float(acc[g * 4 + 3]) + bb[g].w + rr[g].w;
          sum += (v.x + v.y) + (v.z + v.w);
          *(float4*)(xo + c * 32 + g * 4) = v;
          acc[g * 4] = __float_as_uint(v.x); acc[g * 4 + 1] = __float_as_uint(v.y);
          acc[g * 4 + 2] = __float_as_uint(v.z); acc[g * 4 + 3] = __float_as_uint(v.w);
        }
        tc::tmem_st32(taddr + c * 32, acc);
#pragma unroll
        for (int g = 0; g < 8; ++g) rr[g] = rn[g];
      }
      const int row = q * 32 + lane;
      red[half * 128 + row] = sum;
      tc::tmem_st_wait();
      tc::named_bar_sync(1 + q, 64);
      const float mean = (sum + red[(half ^ 1) * 128 + row]) * (1.f / RBN);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t acc[32];
        tc::tmem_ld32(taddr + c * 32, acc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = __uint_as_float(acc[i]) - mean;
          var = fmaf(d, d, var);
        }
      }
      red[256 + half * 128 + row] = var;
      tc::named_bar_sync(1 + q, 64);
      const float rstd = rsqrtf((var + red[256 + (half ^ 1) * 128 + row]) * (1.f / RBN) + p.eps);
      const float* gam = p.gamma + half * 128;
      const float* bet = p.beta + half * 128;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t acc[32];
        tc::tmem_ld32(taddr + c * 32, acc);
        float4 gg[8], be[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          gg[g] = __ldg((const float4*)(gam + c * 32 + g * 4));
          be[g] = __ldg((const float4*)(bet + c * 32 + g * 4));
        }
        tc::tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          y[g * 4] = (__uint_as_float(acc[g * 4]) - mean) * rstd * gg[g].x + be[g].x;
          y[g * 4 + 1] = (__uint_as_float(acc[g * 4 + 1]) - mean) * rstd * gg[g].y + be[g].y;
          y[g * 4 + 2] = (__uint_as_float(acc[g * 4 + 2]) - mean) * rstd * gg[g].z + be[g].z;
          y[g * 4 + 3] = (__uint_as_float(acc[g * 4 + 3]) - mean) * rstd * gg[g].w + be[g].w;
        }
        if (p.t_f32) {
          float* to = (float*)p.t_out + (long)r * p.ldt + half * 128 + c * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g) *(float4*)(to + g * 4) = make_float4(y[g * 4], y[g * 4 + 1], y[g * 4 + 2], y[g * 4 + 3]);
        } else {
          bf16* to = (bf16*)p.t_out + (long)r * p.ldt + half * 128 + c * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            uint32_t* vv = (uint32_t*)&v;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 hh = __floats2bfloat162_rn(y[g * 8 + 2 * e], y[g * 8 + 2 * e + 1]);
              vv[e] = *(uint32_t*)&hh;
            }
            *(uint4*)(to + g * 8) = v;
          }
        }
      }
    } else {
      const bool rope = (int)blockIdx.y < p.rope_tiles;
      const int pos = (r % p.L) % p.rope_len;
      const float* ct = p.cos_t + (long)pos * (RBN / 2) + half * 64;
      const float* stb = p.sin_t + (long)pos * (RBN / 2) + half * 64;
      const float* bias = p.bias ? p.bias + n0 + half * 128 : nullptr;
      bf16* to = (bf16*)p.t_out + (long)r * p.ldt + n0 + half * 128;
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t acc[32];
        tc::tmem_ld32(taddr + c * 32, acc);
        float4 bb[8], cc[4], ss[4];
#pragma unroll
        for (int g = 0; g < 8; ++g) bb[g] = bias ? __ldg((const float4*)(bias + c * 32 + g * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (rope) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            cc[g] = *(const float4*)(ct + c * 16 + g * 4);
            ss[g] = *(const float4*)(stb + c * 16 + g * 4);
          }
        }
        tc::tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          // the projection's own bf16 rounding first (what autocast hands to apply_rotary_enc), rotation in fp32
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(acc[g * 4]) + bb[g].x, __uint_as_float(acc[g * 4 + 1]) + bb[g].y);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(acc[g * 4 + 2]) + bb[g].z, __uint_as_float(acc[g * 4 + 3]) + bb[g].w);
          if (rope) {
            const float4 c4 = cc[g >> 1], s4 = ss[g >> 1];
            const float c0 = (g & 1) ? c4.z : c4.x, c1 = (g & 1) ? c4.w : c4.y;
            const float s0 = (g & 1) ? s4.z : s4.x, s1 = (g & 1) ? s4.w : s4.y;
            const float2 a = __bfloat1622float2(h0), b2 = __bfloat1622float2(h1);
            h0 = __floats2bfloat162_rn(a.x * c0 - a.y * s0, a.x * s0 + a.y * c0);
            h1 = __floats2bfloat162_rn(b2.x * c1 - b2.y * s1, b2.x * s1 + b2.y * c1);
          }
          w[g * 2] = *(uint32_t*)&h0;
          w[g * 2 + 1] = *(uint32_t*)&h1;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) *(uint4*)(to + c * 32 + g * 8) = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, RBN);
  }
}

int launch_row(int mode, const void* A, long lda, const void* W, RowP& p, cudaStream_t st) {
  MS2_CHECK_ARG(A && W && ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && lda % 8 == 0, "gemm_row: operand alignment");
  MS2_CHECK_ARG(p.M > 0 && p.M % RBM == 0 && p.N % RBN == 0 && p.K % 8 == 0 && p.K >= 16,
                "gemm_row: M must be a multiple of 128, N of 256, K of 8 (got %d, %d, %d)", p.M, p.N, p.K);
  MS2_CHECK_ARG(!p.bias || (uintptr_t)p.bias % 16 == 0, "gemm_row: bias alignment");
  p.num_kb = (p.K + RBK - 1) / RBK;
  CUtensorMap tmA, tmW;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M}, str[1] = {(uint64_t)lda};
    const uint32_t box[2] = {RBK, RBM};
    if ((rc = tc::make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N}, str[1] = {(uint64_t)p.K};
    const uint32_t box[2] = {RBK, RBN};
    if ((rc = tc::make_tmap_bf16(&tmW, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  dim3 grid(p.M / RBM, p.N / RBN);
  if (mode == 1) {
    static bool attr = false;
    if (!attr) {
      MS2_CUDA(cudaFuncSetAttribute(gemm_row_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM), "gemm_row attr");
      attr = true;
    }
    ms2_launch(gemm_row_kernel<1>, grid, R_THREADS, R_SMEM, st, tmA, tmW, p);
  } else {
    static bool attr = false;
    if (!attr) {
      MS2_CUDA(cudaFuncSetAttribute(gemm_row_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM), "gemm_row attr");
      attr = true;
    }
    ms2_launch(gemm_row_kernel<0>, grid, R_THREADS, R_SMEM, st, tmA, tmW, p);
  }
  MS2_CHECK_LAUNCH("gemm_row_kernel");
  return MS2_OK;
}

}  // namespace

extern "C" int ms2_gemm_res_ln(const void* A, long lda, const void* W, const float* bias, const float* residual, long ldr,
                               float* x_out, long ldx, const float* gamma, const float* beta, float eps, void* t_out, int t_dt,
                               long ldt, int M, int K, void* stream) {
  MS2_CHECK_ARG(residual && x_out && gamma && beta && t_out, "gemm_res_ln: null pointer");
  MS2_CHECK_ARG(t_dt == MS2_F32 || t_dt == MS2_BF16, "gemm_res_ln: LayerNorm output must be fp32 or bf16");
  MS2_CHECK_ARG(((uintptr_t)residual % 16 == 0) && ((uintptr_t)x_out % 16 == 0) && ((uintptr_t)t_out % 16 == 0) &&
                    ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0) && ldr % 4 == 0 && ldx % 4 == 0 &&
                    ldt % (t_dt == MS2_F32 ? 4 : 8) == 0,
                "gemm_res_ln: 16-byte alignment of rows");
  MS2_CHECK_ARG((const void*)t_out != (const void*)x_out && (const void*)t_out != (const void*)residual && t_out != A,
                "gemm_res_ln: the LayerNorm output must not alias an input");
  RowP p = {};
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.x_out = x_out; p.ldx = ldx; p.t_out = t_out; p.ldt = ldt;
  p.t_f32 = t_dt == MS2_F32; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.M = M; p.N = RBN; p.K = K; p.L = M; p.rope_len = 1;
  return launch_row(1, A, lda, W, p, (cudaStream_t)stream);
}

extern "C" int ms2_gemm_rope(const void* A, long lda, const void* W, const float* bias, void* out, long ldo, int M, int N,
                             int K, int L, int rope_cols, const float* cos_t, const float* sin_t, int table_len,
                             void* stream) {
  MS2_CHECK_ARG(out && ((uintptr_t)out % 16 == 0) && ldo % 8 == 0 && out != A, "gemm_rope: output alignment / aliasing");
  MS2_CHECK_ARG(rope_cols >= 0 && rope_cols <= N && rope_cols % RBN == 0, "gemm_rope: rope_cols must be a multiple of 256");
  MS2_CHECK_ARG(L > 0 && M % L == 0, "gemm_rope: M must be a multiple of the rows per item");
  MS2_CHECK_ARG(rope_cols == 0 || (cos_t && sin_t && table_len > 0 && ((uintptr_t)cos_t % 16 == 0) && ((uintptr_t)sin_t % 16 == 0)),
                "gemm_rope: rotation tables");
  RowP p = {};
  p.bias = bias; p.t_out = out; p.ldt = ldo; p.cos_t = cos_t; p.sin_t = sin_t; p.rope_len = table_len > 0 ? table_len : 1;
  p.rope_tiles = rope_cols / RBN; p.L = L;
  p.M = M; p.N = N; p.K = K;
  return launch_row(0, A, lda, W, p, (cudaStream_t)stream);
}
