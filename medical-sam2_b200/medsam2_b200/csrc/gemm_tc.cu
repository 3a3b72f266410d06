// tcgen05 / TMEM / TMA bf16 GEMM with the fused epilogue of ms2_gemm (include/medsam2_b200.h):
//     out[M,N] = residual + colscale * act(A[M,K] @ W[N,K]^T + bias)
// Replaces the cuBLAS addmm / 1x1-conv calls behind every nn.Linear of the hot path (SURVEY §8(a)
// a2,a3,a6,a8,a9).  Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer  : A (128 x 64) and W (BN x 64) boxes, 128B-swizzled, ring of `stages`
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into one of two
//                               TMEM accumulator stages; tcgen05.commit releases smem slots / signals epilogue
//   warps 2..5  epilogue      : tcgen05.ld 32 lanes x 32 columns -> bias / act / colscale / residual ->
//                               16-byte global stores (fp32 or bf16), overlapping the next tile's main loop
// Both operands are K-major, so no transposes exist anywhere; M/N/K tails are handled by TMA
// out-of-bounds zero fill and masked stores.  BN (32..256) is a run-time choice per problem shape.
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64, MAX_STAGES = 8, ACC_STAGES = 2, TMEM_COLS = 512;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int NUM_THREADS = 192;

struct GemmP {
  const float* bias;
  const float* colscale;
  const float* residual;
  long ldr;
  void* out;
  long ldo;
  int o_dt, M, N, K, act, BN, tiles_n, tiles, num_kb, stages;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return fmaxf(v, 0.f);
  if (act == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint32_t* tmem_ptr = (uint32_t*)(acc_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * p.BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          tc::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
          tc::tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
          tc::tma_load_2d(sa + A_STAGE_BYTES, &tmW, &full[stage], kb * BK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(BM, p.BN, 0, 0);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      tc::mbar_wait(&acc_empty[as], aphase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            tc::umma_bf16(d_tmem, tc::desc_kmajor_sw128(sa + k * 32), tc::desc_kmajor_sw128(sb + k * 32), idesc,
                          (kb | k) ? 1u : 0u);
          }
          tc::umma_commit(&empty[stage]);
          if (kb == p.num_kb - 1) tc::umma_commit(&acc_full[as]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * p.BN;
      const int row = m0 + q * 32 + lane;
      tc::mbar_wait(&acc_full[as], aphase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + c, r);
        tc::tmem_ld_wait();
        const int nb = n0 + c;
        if (row < p.M && nb < p.N) {
          const float* res = p.residual ? p.residual + (long)row * p.ldr + nb : nullptr;
#pragma unroll
          for (int g = 0; g < 32; g += 8) {
            if (nb + g >= p.N) continue;
            const bool second = nb + g + 4 < p.N;          // N % 4 == 0: groups of 4 are all-or-nothing
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g + j]);
            if (p.bias) {
              const float4 b0 = __ldg((const float4*)(p.bias + nb + g));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              if (second) {
                const float4 b1 = __ldg((const float4*)(p.bias + nb + g + 4));
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
            }
            if (p.act) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act);
            }
            if (p.colscale) {
              const float4 s0 = __ldg((const float4*)(p.colscale + nb + g));
              v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
              if (second) {
                const float4 s1 = __ldg((const float4*)(p.colscale + nb + g + 4));
                v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
              }
            }
            if (res) {
              const float4 r0 = *(const float4*)(res + g);
              v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
              if (second) {
                const float4 r1 = *(const float4*)(res + g + 4);
                v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
            }
            if (p.o_dt == MS2_F32) {
              float* o = (float*)p.out + (long)row * p.ldo + nb + g;
              *(float4*)o = make_float4(v[0], v[1], v[2], v[3]);
              if (second) *(float4*)(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              bf16* o = (bf16*)p.out + (long)row * p.ldo + nb + g;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
              if (second) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                uint4 u;
                u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
                *(uint4*)o = u;
              } else {
                uint2 u;
                u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1;
                *(uint2*)o = u;
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int pick_bn(int M, int N, int K) {
  const int sms = tc::sm_count();
  const long tm = (M + BM - 1) / BM;
  const long kb = (K + BK - 1) / BK;
  int best = 32;
  double best_cost = 1e30;
  const int nmax = ((N + 31) / 32) * 32;
  for (int bn = 32; bn <= 256 && bn <= (nmax < 32 ? 32 : nmax); bn += 32) {
    const long tn = (N + bn - 1) / bn;
    const long waves = (tm * tn + sms - 1) / sms;
    const double per_tile = kb * (390.0 + 3.05 * bn) + 200.0 + 8.0 * bn;
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

bool ms2_gemm_tc_supported(int a_dt, int w_dt, long lda, long ldo, int M, int N, int K) {
  return a_dt == MS2_BF16 && w_dt == MS2_BF16 && M >= 64 && K % 8 == 0 && K >= 16 && lda % 8 == 0 && N % 8 == 0 &&
         ldo % 8 == 0;
}

int ms2_gemm_tc_launch(const void* A, long lda, const void* W, const float* bias, const float* colscale,
                       const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K,
                       int act, cudaStream_t st) {
  MS2_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)out % 16 == 0),
                "gemm_tc: operands must be 16-byte aligned");
  MS2_CHECK_ARG(!residual || (((uintptr_t)residual % 16 == 0) && ldr % 4 == 0), "gemm_tc: residual alignment");
  MS2_CHECK_ARG(!bias || (uintptr_t)bias % 16 == 0, "gemm_tc: bias alignment");
  MS2_CHECK_ARG(!colscale || (uintptr_t)colscale % 16 == 0, "gemm_tc: colscale alignment");
  MS2_CHECK_ARG(o_dt == MS2_F32 || o_dt == MS2_BF16, "gemm_tc: bad output dtype");
  GemmP p;
  p.bias = bias; p.colscale = colscale; p.residual = residual; p.ldr = ldr; p.out = out; p.ldo = ldo; p.o_dt = o_dt;
  p.M = M; p.N = N; p.K = K; p.act = act;
  p.BN = pick_bn(M, N, K);
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.tiles = ((M + BM - 1) / BM) * p.tiles_n;
  p.num_kb = (K + BK - 1) / BK;
  const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
  p.stages = SMEM_BUDGET / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;

  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)lda};
    const uint32_t box[2] = {BK, BM};
    int rc = tc::make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K};
    const uint32_t box[2] = {BK, (uint32_t)p.BN};
    int rc = tc::make_tmap_bf16(&tmW, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MS2_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 2048),
             "gemm_tc attr");
    attr_set = true;
  }
  const int grid = p.tiles < tc::sm_count() ? p.tiles : tc::sm_count();
  gemm_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmW, p);
  MS2_CHECK_LAUNCH("gemm_tc_kernel");
  return MS2_OK;
}
