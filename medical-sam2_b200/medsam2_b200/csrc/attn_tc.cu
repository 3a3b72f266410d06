// tcgen05 / TMEM / TMA flash attention (no mask), bf16 operands, fp32 softmax statistics:
//     O = softmax(scale * Q K^T) V          Q [B,Hh,Lq,D], K/V [B,Hh,Lk,D] through element strides
// Replaces F.scaled_dot_product_attention at transformer.py:252-258,318 (memory self/cross attention,
// one head of 256) and hieradet.py:72-76 (Hiera global blocks, heads of 96).  SURVEY §8(a) a2/a6.
//
// One CTA owns a 128-query tile and streams a contiguous range of key tiles (split-KV over gridDim.z so
// that 32 query tiles still fill 148 SMs; partial (O, m, l) are merged by attn_combine_kernel):
//   warp 0      TMA producer : Q once; K and V tiles (BKV keys) through two 2-deep rings, 128B-swizzled
//   warp 1      MMA issuer   : S = Q K^T   (M=128, N=BKV, K=D; Q is copied ONCE from smem into TENSOR MEMORY and
//                              used as the A operand from there: with both operands in smem the 128x16 Q slice
//                              would be re-read for every instruction and the ~56 B/clk operand fetch, not the
//                              tensor pipe, would set the pace)                                  -> TMEM S[j&1]
//                              O += P V    (M=128, N=D, K=BKV, P from TENSOR MEMORY, V MN-major smem) -> TMEM O
//   warps 2..9  softmax      : two warps per TMEM lane quarter split the S columns of their 32 query rows;
//                              thread = (row, column half): tcgen05.ld of its S slice, online max/sum in the log2
//                              domain with LAZY rescaling (O is only rescaled in TMEM when the running max
//                              grows by > 2^8), P -> bf16 pairs -> tcgen05.st over the S columns it came from
//                              (no shared-memory round trip, no proxy fence); final O/l -> global
// QK^T of tile j+1 is issued before P V of tile j, so the tensor pipe works under the softmax.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BQ = 128;
constexpr int NUM_THREADS = 320;
constexpr float LAZY_TAU = 8.0f;

struct AttnTcP {
  void* o;
  long o_bs, o_hs, o_ts;
  float* opart;   // [nsplit][B*Hh][Lq][D] fp32 un-normalised partial outputs (nsplit > 1)
  float* ml;      // [nsplit][B*Hh][Lq][2]  (running max in log2 units, row sum)
  int B, Hh, Lq, Lk;
  float c;        // scale * log2(e)
  int ntiles, tiles_per_split, nsplit;
  float tau;      // lazy-rescale threshold (log2 units)
  int force_part; // write the un-normalised partial (opart, ml) even when nsplit == 1 (split-KV across GPUs)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax
// polynomial (max relative error 7.6e-5, far below the bf16 rounding of P, 3.9e-3), 2^n by adding n to the exponent
// bits.  The 1.5*2^23 constant leaves n in the low mantissa bits of t.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.05520550534f, f, 0.24261397123f);
  p = fmaf(p, f, 0.69325476885f);
  p = fmaf(p, f, 0.99992769957f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// packed fp32 pairs (Blackwell FFMA2 / FADD2): one issue slot for two elements of the softmax inner loop
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%5, %5};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{\n.reg .b64 ra, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rd, {%0, %1};\nadd.rn.f32x2 rd, rd, ra;\nmov.b64 {%0, %1}, rd;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}

// D = head dim of Q/K, DV = head dim of V/O (DV < D: attention over un-projected 64-d memory values, the value
// projection is applied to the 64-d result afterwards - softmax rows sum to 1), BKV = keys per
// tile, KST = depth of the K ring (V ring: 2).
template <int D, int DV, int BKV, int KST>
struct Cfg {
  static constexpr int DCH = (D + 63) / 64;             // 64-element (128 B) column chunks of Q/K
  static constexpr int DCHV = (DV + 63) / 64;           // ... of V
  static constexpr int Q_BYTES = DCH * BQ * 128;
  static constexpr int K_BYTES = DCH * BKV * 128;       // one K stage
  static constexpr int V_BYTES = DCHV * BKV * 128;      // one V stage
  static constexpr int SMEM = Q_BYTES + KST * K_BYTES + 2 * V_BYTES + 1024 /*align*/ + 4096 /*barriers, max/sum exchange*/;
  static constexpr int S_COL = 256;                     // TMEM: O at [0,DV), S stages at 256 + st*BKV
  static constexpr int Q_COL = D == 256 ? 384 : 128;    // Q (A operand of S = Q K^T) resident in TMEM, D/2 columns
};

template <int D, int DV, int BKV, int KST, int POLY>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcP p) {
  using C = Cfg<D, DV, BKV, KST>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;
  uint8_t* sV = sK + KST * C::K_BYTES;
  uint64_t* bars = (uint64_t*)(sV + 2 * C::V_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // <= 4
  uint64_t* k_empty = bars + 5;       // <= 4
  uint64_t* v_full = bars + 9;        // 2
  uint64_t* v_empty = bars + 11;      // 2
  uint64_t* s_full = bars + 13;       // 2
  uint64_t* p_full = bars + 15;       // 1
  uint64_t* o_ready = bars + 16;      // 1
  uint64_t* q_tmem = bars + 17;       // 1: Q copied into tensor memory
  uint32_t* tmem_ptr = (uint32_t*)(bars + 18);
  float* mxbuf = (float*)(bars + 19);    // [tile parity][half][128 rows] row-max exchange
  float* lbuf = mxbuf + 512;             // [half][128 rows] row-sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y, b = bh / p.Hh, h = bh - b * p.Hh;
  const int split = blockIdx.z;
  const int t_begin = split * p.tiles_per_split;
  int n = p.ntiles - t_begin;
  if (n > p.tiles_per_split) n = p.tiles_per_split;     // >= 1 by construction on the host

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) {
      tc::mbar_init(&k_full[s], 1);
      tc::mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&v_full[s], 1);
      tc::mbar_init(&v_empty[s], 1);
      tc::mbar_init(&s_full[s], 1);
    }
    tc::mbar_init(p_full, 8);
    tc::mbar_init(o_ready, 1);
    tc::mbar_init(q_tmem, 8);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();      // barriers, tensor-memory allocation and descriptor prefetch above overlap the preceding kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < C::DCH; ++c) tc::tma_load_4d(sQ + c * BQ * 128, &tmQ, q_full, c * 64, q0, h, b);
      // K runs KST-1 tiles ahead of V: the K ring is deeper, and S = Q K^T of tile j+1 is issued before P V of j
      auto load_k = [&](int j) {
        const int ks = j % KST;
        tc::mbar_wait(&k_empty[ks], ((uint32_t)(j / KST) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&k_full[ks], C::K_BYTES);
#pragma unroll
        for (int c = 0; c < C::DCH; ++c)
          tc::tma_load_4d(sK + ks * C::K_BYTES + c * BKV * 128, &tmK, &k_full[ks], c * 64, (t_begin + j) * BKV, h, b);
      };
      for (int j = 0; j < KST - 1 && j < n; ++j) load_k(j);
      for (int j = 0; j < n; ++j) {
        if (j + KST - 1 < n) load_k(j + KST - 1);
        const int st = j & 1;
        tc::mbar_wait(&v_empty[st], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&v_full[st], C::V_BYTES);
#pragma unroll
        for (int c = 0; c < C::DCHV; ++c)
          tc::tma_load_4d(sV + st * C::V_BYTES + c * BKV * 128, &tmV, &v_full[st], c * 64, (t_begin + j) * BKV, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = tc::make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, DV, 0, 1);
    const uint32_t aK = tc::smem_u32(sK), aV = tc::smem_u32(sV);
    const uint32_t tO = tmem_base, tS = tmem_base + C::S_COL, tQ = tmem_base + C::Q_COL;

    auto issue_qk = [&](int j) {
      const int st = j & 1, ks = j % KST;
      // S[st] still holds P of tile j-2 in its first BKV/2 columns: P V of tile j-2 must have finished READING it
      // before this S = Q K^T starts writing (back-to-back independent MMAs overlap in the tensor pipe, so issue
      // order alone does not protect the operand read)
      if (j >= 2) tc::mbar_wait(o_ready, (uint32_t)(j - 2) & 1u);
      tc::mbar_wait(&k_full[ks], (uint32_t)(j / KST) & 1u);
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t db = tc::desc_kmajor_sw128(aK + ks * C::K_BYTES + (kk >> 2) * BKV * 128 + (kk & 3) * 32);
          tc::umma_bf16_ts(tS + st * BKV, tQ + kk * 8, db, idesc_qk, kk ? 1u : 0u);   // 16 dims = 8 columns of Q
        }
        tc::umma_commit(&k_empty[ks]);
        tc::umma_commit(&s_full[st]);
      }
      __syncwarp();
    };

    tc::mbar_wait(q_tmem, 0);
    tc::tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < n; ++j) {
      if (j + 1 < n) issue_qk(j + 1);
      const int st = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1u;
      tc::mbar_wait(&v_full[st], ph);
      tc::mbar_wait(p_full, (uint32_t)j & 1u);
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t db = tc::desc_mnmajor_sw128(aV + st * C::V_BYTES + kk * 2048, BKV * 128);
          tc::umma_bf16_ts(tO, tS + st * BKV + kk * 8, db, idesc_pv, (j | kk) ? 1u : 0u);   // 16 keys = 8 columns of P
        }
        tc::umma_commit(&v_empty[st]);
        tc::umma_commit(o_ready);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..9) =====================
    // Two warps share each TMEM lane quarter (= 32 query rows) and split the S columns of a tile, so every SM
    // sub-partition has two softmax warps to overlap MUFU/FMA latency; the row maximum is exchanged through shared
    // memory with a 64-thread named barrier, the row sums are merged once at the end.
    constexpr int HC = BKV / 2;                     // S columns per warp
    const int qtr = warp & 3, half = (warp - 2) >> 2;
    const int row = qtr * 32 + lane;                // query row inside the tile
    const uint32_t lane_addr = (uint32_t)(qtr * 32) << 16;
    const uint32_t tO = tmem_base + lane_addr, tS = tmem_base + lane_addr + C::S_COL;
    float m_used = 0.f, l = 0.f;
    const int last_valid = p.Lk - (p.ntiles - 1) * BKV;   // valid keys of the globally last tile

    // ---- Q: swizzled smem tile (TMA) -> tensor memory, row = lane, two bf16 per 32-bit column; the two warps of a
    //      lane quarter take alternate 16-byte pieces-groups (32-column blocks of the packed row)
    {
      tc::mbar_wait(q_full, 0);
      const uint32_t aQ = tc::smem_u32(sQ);
      constexpr int QW = D / 2;                         // packed words per row
#pragma unroll
      for (int blk = 0; blk < (QW + 15) / 16; ++blk) {  // 16 words = 32 bf16 = 4 pieces of 16 bytes
        if ((blk & 1) != half) continue;
        const int ch = blk >> 1, g0 = (blk & 1) * 4;    // 64-column chunk, first piece inside it
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 v = tc::lds128(aQ + ch * BQ * 128 + row * 128 + (((g0 + g) ^ (row & 7)) << 4));
          w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
        }
        tc::tmem_st16(tmem_base + lane_addr + C::Q_COL + blk * 16, w);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_tmem);
    }

    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      tc::mbar_wait(&s_full[st], (uint32_t)(j >> 1) & 1u);
      tc::tc_fence_after();
      uint32_t r[HC / 32][32];
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) tc::tmem_ld32(tS + st * BKV + half * HC + c * 32, r[c]);
      tc::tmem_ld_wait();

      float mx = -INFINITY;
      if ((t_begin + j == p.ntiles - 1) && (last_valid < BKV)) {
#pragma unroll
        for (int c = 0; c < HC / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (half * HC + c * 32 + i >= last_valid) r[c][i] = 0xff800000u;     // -inf: key beyond Lk
      }
#pragma unroll
      for (int c = 0; c < HC / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[c][i]));
      mx *= p.c;                                      // p.c > 0: max commutes with the scaling
      float* mslot = mxbuf + (j & 1) * 256;
      mslot[half * 128 + row] = mx;
      tc::named_bar_sync(1 + qtr, 64);
      mx = fmaxf(mx, mslot[(half ^ 1) * 128 + row]);
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + p.tau;
        if (__any_sync(0xffffffffu, need)) {
          tc::mbar_wait(o_ready, (uint32_t)(j - 1) & 1u);      // P V of tile j-1 has landed in O
          tc::tc_fence_after();
          float alpha = 1.f;
          if (need) {
            alpha = ex2(m_used - mx);
            m_used = mx;
            l *= alpha;
          }
#pragma unroll 1
          for (int c = half; c < DV / 32; c += 2) {             // the two warps of a quarter split the O columns
            uint32_t o[32];
            tc::tmem_ld32(tO + c * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st32(tO + c * 32, o);
          }
          tc::tmem_st_wait();
        }
      }
      // p = 2^(s*c - m_used), row sum in fp32, pack to bf16 pairs
      uint32_t pk[HC / 2];
      const float nm = -m_used;
      // D = 96 (Hiera global blocks): per 128x128 tile the tensor pipe needs 768 cycles, the 16384 exponentials 1024
      // MUFU cycles -> a share of them (p.poly eighths, 0..4) is evaluated by ex2_poly on the FMA pipe instead
#pragma unroll
      for (int c = 0; c < HC / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = fmaf(__uint_as_float(r[c][i]), p.c, nm);
          const float x1 = fmaf(__uint_as_float(r[c][i + 1]), p.c, nm);
          const int e = i & 7;
          const bool poly0 = D == 96 && ((POLY >= 1 && e == 2) || (POLY >= 3 && e == 6));
          const bool poly1 = D == 96 && ((POLY >= 2 && e == 4) || (POLY >= 4 && e == 0));
          const float p0 = poly0 ? ex2_poly(x0) : ex2(x0);
          const float p1 = poly1 ? ex2_poly(x1) : ex2(x1);
          l += p0 + p1;
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[(c * 32 + i) >> 1] = *(uint32_t*)&hh;
        }
      // Every warp observes EVERY phase of o_ready (one per tile): an mbarrier only carries a phase parity, and a
      // waiter that skips phases (the rescale above is rare) can no longer tell phase j-1 from j-3.  P V of tile
      // j-1 was issued a whole softmax ago, so this wait is normally already satisfied.
      if (j > 0) tc::mbar_wait(o_ready, (uint32_t)(j - 1) & 1u);
      // P overwrites the first BKV/2 columns of this S stage: row = lane, two bf16 keys per 32-bit column
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) tc::tmem_st16(tS + st * BKV + half * (HC / 2) + c * 16, &pk[c * 16]);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
    }

    // ---- epilogue: merge the two partial row sums, then each warp writes its share of the O columns
    lbuf[half * 128 + row] = l;
    tc::named_bar_sync(1 + qtr, 64);
    l += lbuf[(half ^ 1) * 128 + row];
    tc::mbar_wait(o_ready, (uint32_t)(n - 1) & 1u);
    tc::tc_fence_after();
    const int qi = q0 + row;
    if (p.nsplit == 1 && !p.force_part) {
      const float inv = 1.f / l;
      bf16* orow = (bf16*)p.o + (long)b * p.o_bs + (long)h * p.o_hs + (long)qi * p.o_ts;
#pragma unroll 1
      for (int c = half; c < DV / 32; c += 2) {
        uint32_t o[32];
        tc::tmem_ld32(tO + c * 32, o);
        tc::tmem_ld_wait();
        if (qi < p.Lq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            uint32_t* vv = (uint32_t*)&v;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv,
                                                        __uint_as_float(o[g * 8 + 2 * i + 1]) * inv);
              vv[i] = *(uint32_t*)&hh;
            }
            *(uint4*)(orow + c * 32 + g * 8) = v;
          }
        }
      }
    } else {
      const long rix = ((long)split * gridDim.y + bh) * p.Lq + qi;
      float* orow = p.opart + rix * DV;
#pragma unroll 1
      for (int c = half; c < DV / 32; c += 2) {
        uint32_t o[32];
        tc::tmem_ld32(tO + c * 32, o);
        tc::tmem_ld_wait();
        if (qi < p.Lq) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *(float4*)(orow + c * 32 + g * 4) =
                make_float4(__uint_as_float(o[g * 4]), __uint_as_float(o[g * 4 + 1]), __uint_as_float(o[g * 4 + 2]),
                            __uint_as_float(o[g * 4 + 3]));
        }
      }
      if (qi < p.Lq && half == 0) *(float2*)(p.ml + rix * 2) = make_float2(m_used, l);
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Memory cross-attention, TWO query tiles per CTA (D = 256 keys, DV = 64 un-projected values, 64-key tiles).
// With one 128-query tile per CTA the 32 query tiles of a frame each stream the whole K bank: ~6 TB/s of L2->SM
// traffic at 0.96 PFLOP/s, i.e. the L2, not the tensor pipe, sets the pace (profiles/r1_attn_dv_ncu.txt).  Here a
// CTA owns 256 queries as two independent 128-row streams a/b that share every K/V tile (half the operand traffic
// per FLOP) and ping-pong on the tensor pipe: while one stream's softmax runs, the other stream's S = Q K^T or
// O += P V executes.
//   warp 0      TMA producer : Q_a, Q_b (through one 64 KB staging buffer), K ring (4 x 32 KB), V ring (2 x 8 KB)
//   warp 1      MMA issuer   : order  P V_a(j), S_a(j+1), P V_b(j), S_b(j+1)  (Q and P operands in tensor memory)
//   warps 2..5  softmax of stream a (one warp per TMEM lane quarter, a thread owns a full 64-column S row)
//   warps 6..9  softmax of stream b
// TMEM: O_a [0,64) O_b [64,128) S_a [128,192) S_b [192,256) Q_a [256,384) Q_b [384,512); P overwrites S in place,
// so S_x(j+1) is only issued after P V_x(j) has completed (o_ready_x), and every softmax warp observes every phase.
constexpr int T2_THREADS = 320, T2_KST = 4, T2_BKV = 64;
constexpr bool MC_DEFAULT_ON = true;
constexpr int ATTN_D96_POLY = 2;
constexpr int T2_Q_BYTES = 4 * BQ * 128, T2_K_BYTES = 4 * T2_BKV * 128, T2_V_BYTES = T2_BKV * 128;
constexpr int T2_SMEM = T2_Q_BYTES + T2_KST * T2_K_BYTES + 2 * T2_V_BYTES + 1024 + 1024;

// VAR: bit 0 = independent partial maxima / row sums (breaks the 32-deep FMNMX3 and FADD dependency chains of a row);
//      bits 1..2 = share of the exponentials evaluated by ex2_poly on the FMA pipe instead of MUFU.EX2: 0, 1/4, 3/8, 1/2
template <int VAR>
__global__ void __launch_bounds__(T2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnTcP p) {
  constexpr int BKV = T2_BKV, KST = T2_KST, DV = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T2_Q_BYTES;
  uint8_t* sV = sK + KST * T2_K_BYTES;
  uint64_t* bars = (uint64_t*)(sV + 2 * T2_V_BYTES);
  uint64_t* q_full = bars;            // 2 (stream a, b)
  uint64_t* q_tmem = bars + 2;        // 2: Q_x copied into tensor memory (4 warps each)
  uint64_t* k_full = bars + 4;        // 4
  uint64_t* k_empty = bars + 8;       // 4
  uint64_t* v_full = bars + 12;       // 2
  uint64_t* v_empty = bars + 14;      // 2
  uint64_t* s_full = bars + 16;       // 2 (stream)
  uint64_t* p_full = bars + 18;       // 2 (stream), 4 warps each
  uint64_t* o_ready = bars + 20;      // 2 (stream)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int bh = blockIdx.y, b = bh / p.Hh, h = bh - b * p.Hh;
  const int split = blockIdx.z;
  const int t_begin = split * p.tiles_per_split;
  int n = p.ntiles - t_begin;
  if (n > p.tiles_per_split) n = p.tiles_per_split;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&q_full[s], 1);
      tc::mbar_init(&q_tmem[s], 4);
      tc::mbar_init(&v_full[s], 1);
      tc::mbar_init(&v_empty[s], 1);
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&p_full[s], 4);
      tc::mbar_init(&o_ready[s], 1);
    }
    for (int s = 0; s < KST; ++s) {
      tc::mbar_init(&k_full[s], 1);
      tc::mbar_init(&k_empty[s], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();      // barriers, tensor-memory allocation and descriptor prefetch above overlap the preceding kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&q_full[0], T2_Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c) tc::tma_load_4d(sQ + c * BQ * 128, &tmQ, &q_full[0], c * 64, q0, h, b);
      auto load_k = [&](int j) {
        const int ks = j % KST;
        tc::mbar_wait(&k_empty[ks], ((uint32_t)(j / KST) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&k_full[ks], T2_K_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tc::tma_load_4d(sK + ks * T2_K_BYTES + c * BKV * 128, &tmK, &k_full[ks], c * 64, (t_begin + j) * BKV, h, b);
      };
      for (int j = 0; j < KST - 1 && j < n; ++j) load_k(j);
      // Q_b reuses the staging buffer once stream a's warps have moved Q_a into tensor memory
      tc::mbar_wait(&q_tmem[0], 0);
      tc::mbar_arrive_expect_tx(&q_full[1], T2_Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c) tc::tma_load_4d(sQ + c * BQ * 128, &tmQ, &q_full[1], c * 64, q0 + BQ, h, b);
      for (int j = 0; j < n; ++j) {
        if (j + KST - 1 < n) load_k(j + KST - 1);
        const int st = j & 1;
        tc::mbar_wait(&v_empty[st], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(&v_full[st], T2_V_BYTES);
        tc::tma_load_4d(sV + st * T2_V_BYTES, &tmV, &v_full[st], 0, (t_begin + j) * BKV, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = tc::make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, DV, 0, 1);
    const uint32_t aK = tc::smem_u32(sK), aV = tc::smem_u32(sV);
    auto issue_qk = [&](int x, int j) {          // stream x, tile j: S_x = Q_x K_j^T
      const int ks = j % KST;
      if (x == 0) tc::mbar_wait(&k_full[ks], (uint32_t)(j / KST) & 1u);
      if (j >= 1) tc::mbar_wait(&o_ready[x], (uint32_t)(j - 1) & 1u);     // P V_x(j-1) has finished reading P_x
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t tS = tmem_base + 128 + x * 64, tQ = tmem_base + 256 + x * 128;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint64_t db = tc::desc_kmajor_sw128(aK + ks * T2_K_BYTES + (kk >> 2) * BKV * 128 + (kk & 3) * 32);
          tc::umma_bf16_ts(tS, tQ + kk * 8, db, idesc_qk, kk ? 1u : 0u);
        }
        tc::umma_commit(&s_full[x]);
        if (x == 1) tc::umma_commit(&k_empty[ks]);                         // both streams have consumed K_j
      }
      __syncwarp();
    };
    auto issue_pv = [&](int x, int j) {
      const int st = j & 1;
      if (x == 0) tc::mbar_wait(&v_full[st], (uint32_t)(j >> 1) & 1u);
      tc::mbar_wait(&p_full[x], (uint32_t)j & 1u);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t tO = tmem_base + x * 64, tP = tmem_base + 128 + x * 64;
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t db = tc::desc_mnmajor_sw128(aV + st * T2_V_BYTES + kk * 2048, BKV * 128);
          tc::umma_bf16_ts(tO, tP + kk * 8, db, idesc_pv, (j | kk) ? 1u : 0u);
        }
        tc::umma_commit(&o_ready[x]);
        if (x == 1) tc::umma_commit(&v_empty[st]);
      }
      __syncwarp();
    };
    tc::mbar_wait(&q_tmem[0], 0);
    tc::tc_fence_after();
    issue_qk(0, 0);
    tc::mbar_wait(&q_tmem[1], 0);
    tc::tc_fence_after();
    issue_qk(1, 0);
    for (int j = 0; j < n; ++j) {
      issue_pv(0, j);
      if (j + 1 < n) issue_qk(0, j + 1);
      issue_pv(1, j);
      if (j + 1 < n) issue_qk(1, j + 1);
    }
  } else {
    // ===================== softmax / correction / epilogue: warps 2..5 stream a, 6..9 stream b =====================
    const int x = (warp - 2) >> 2;                  // stream
    const int qtr = warp & 3;
    const int row = qtr * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qtr * 32) << 16;
    const uint32_t tO = tmem_base + lane_addr + x * 64, tS = tmem_base + lane_addr + 128 + x * 64;
    const uint32_t tQ = tmem_base + lane_addr + 256 + x * 128;
    float m_used = 0.f, l = 0.f;
    const int last_valid = p.Lk - (p.ntiles - 1) * BKV;

    {  // Q_x: swizzled staging tile -> tensor memory (row = lane, two bf16 per column)
      tc::mbar_wait(&q_full[x], 0);
      const uint32_t aQ = tc::smem_u32(sQ);
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {
        const int ch = blk >> 1, g0 = (blk & 1) * 4;
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 v = tc::lds128(aQ + ch * BQ * 128 + row * 128 + (((g0 + g) ^ (row & 7)) << 4));
          w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
        }
        tc::tmem_st16(tQ + blk * 16, w);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&q_tmem[x]);
    }

    for (int j = 0; j < n; ++j) {
      tc::mbar_wait(&s_full[x], (uint32_t)j & 1u);
      tc::tc_fence_after();
      uint32_t r[2][32];
      tc::tmem_ld32(tS, r[0]);
      tc::tmem_ld32(tS + 32, r[1]);
      tc::tmem_ld_wait();
      if ((t_begin + j == p.ntiles - 1) && (last_valid < BKV)) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= last_valid) r[c][i] = 0xff800000u;
      }
      float mx = -INFINITY;
      if (VAR & 1) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[c][i]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[c][i]));
      }
      mx *= p.c;
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + p.tau;
        if (__any_sync(0xffffffffu, need)) {
          tc::mbar_wait(&o_ready[x], (uint32_t)(j - 1) & 1u);
          tc::tc_fence_after();
          float alpha = 1.f;
          if (need) {
            alpha = ex2(m_used - mx);
            m_used = mx;
            l *= alpha;
          }
#pragma unroll 1
          for (int c = 0; c < DV / 32; ++c) {
            uint32_t o[32];
            tc::tmem_ld32(tO + c * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st32(tO + c * 32, o);
          }
          tc::tmem_st_wait();
        }
      }
      uint32_t pk[32];
      const float nm = -m_used;
      constexpr int PM = (VAR >> 1) & 3;             // polynomial share: 0, 2/8, 3/8, 4/8 of the elements
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = fmaf(__uint_as_float(r[c][i]), p.c, nm);
          const float x1 = fmaf(__uint_as_float(r[c][i + 1]), p.c, nm);
          // element i of every group of 8: which pipe evaluates 2^x (compile-time pattern, interleaved with MUFU ones)
          const int e = i & 7;
          const bool poly0 = (PM == 1 && e == 2) || (PM == 2 && (e == 2 || e == 6)) || (PM == 3 && (e == 2 || e == 6));
          const bool poly1 = (PM == 1 && e == 6) || (PM == 2 && e == 4) || (PM == 3 && (e == 0 || e == 4));
          const float p0 = poly0 ? ex2_poly(x0) : ex2(x0);
          const float p1 = poly1 ? ex2_poly(x1) : ex2(x1);
          if (VAR & 1) l4[(i >> 1) & 3] += p0 + p1;
          else l += p0 + p1;
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[(c * 32 + i) >> 1] = *(uint32_t*)&hh;
        }
      if (VAR & 1) l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      if (j > 0) tc::mbar_wait(&o_ready[x], (uint32_t)(j - 1) & 1u);      // observe every phase (see attn_tc_kernel)
      tc::tmem_st16(tS, &pk[0]);
      tc::tmem_st16(tS + 16, &pk[16]);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[x]);
    }

    tc::mbar_wait(&o_ready[x], (uint32_t)(n - 1) & 1u);
    tc::tc_fence_after();
    const int qi = q0 + x * BQ + row;
    if (p.nsplit == 1 && !p.force_part) {
      const float inv = 1.f / l;
      bf16* orow = (bf16*)p.o + (long)b * p.o_bs + (long)h * p.o_hs + (long)qi * p.o_ts;
#pragma unroll 1
      for (int c = 0; c < DV / 32; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(tO + c * 32, o);
        tc::tmem_ld_wait();
        if (qi < p.Lq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            uint32_t* vv = (uint32_t*)&v;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv,
                                                        __uint_as_float(o[g * 8 + 2 * i + 1]) * inv);
              vv[i] = *(uint32_t*)&hh;
            }
            *(uint4*)(orow + c * 32 + g * 8) = v;
          }
        }
      }
    } else {
      const long rix = ((long)split * gridDim.y + bh) * p.Lq + qi;
      float* orow = p.opart + rix * DV;
#pragma unroll 1
      for (int c = 0; c < DV / 32; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(tO + c * 32, o);
        tc::tmem_ld_wait();
        if (qi < p.Lq) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *(float4*)(orow + c * 32 + g * 4) =
                make_float4(__uint_as_float(o[g * 4]), __uint_as_float(o[g * 4 + 1]), __uint_as_float(o[g * 4 + 2]),
                            __uint_as_float(o[g * 4 + 3]));
        }
      }
      if (qi < p.Lq) *(float2*)(p.ml + rix * 2) = make_float2(m_used, l);
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Memory cross-attention on CTA PAIRS (thread-block cluster of 2, D = 256 keys, DV = 64 values, 128-key tiles).
// attn_tc2 shares every K/V tile between two query tiles of ONE CTA; its price is tensor memory: O_a O_b S_a S_b Q_a Q_b
// fill all 512 columns, so P has to overwrite S in place and S_x(j+1) cannot start before P V_x(j) has finished — the
// per-stream chain S -> softmax -> P V is serial (ncu: tensor pipe 54 % active).  Here the two query tiles live in the
// two CTAs of a cluster and share K/V through TMA MULTICAST: each CTA fetches half of every tile and the hardware
// writes it into both shared memories, so the L2 -> SM operand traffic per FLOP is that of attn_tc2, while each SM
// holds ONE stream with room for a double-buffered S and a separate P:  O [0,64)  P [64,128)  S0 [128,256)  S1 [256,384)
// Q [384,512).  The issue order is S0 S1 | PV0 S2 | PV1 S3 | ...: S(j+2) follows P V(j) without waiting for it (P does not
// live in S), so the pipe never drains between tiles and softmax(j+1) has P V(j) + S(j+2) to hide under; 128-key tiles
// halve the barrier round trips per key.  Empty-slot signals (tcgen05.commit) are multicast to both CTAs: a slot is refilled only when both
// CTAs' tensor pipes have finished reading it.  Q is staged through K slot 1 (one cluster barrier orders that).
#ifdef MS2_ATTN_TRACE
// debug timeline of one CTA (tools/ubench/attn_trace.cu): clock64 of 16 events for tiles 16..47
__device__ long long g_trace[32 * 16];
#define MS2_TRACE(j, ev)                                                                                         \
  do {                                                                                                           \
    if (blockIdx.x == 0 && blockIdx.z == 0 && (j) >= 16 && (j) < 48 && (threadIdx.x & 31) == 0)                    \
      g_trace[((j) - 16) * 16 + (ev)] = clock64();                                                               \
  } while (0)
#if defined(MS2_ATTN_TRACE_ISSUE) || defined(MS2_ATTN_TRACE_SPIN)     // events 9..13 = issue progress of S(j) / idle observers instead of the second softmax warp
#define MS2_TRACE_SM(j, ev) do { if ((ev) < 9) MS2_TRACE(j, ev); } while (0)
#else
#define MS2_TRACE_SM(j, ev) MS2_TRACE(j, ev)
#endif
#else
#define MS2_TRACE(j, ev)
#define MS2_TRACE_SM(j, ev)
#endif
#ifdef MS2_ATTN_TRACE_SPIN
constexpr int MC_BKV = 128, MC_KST = 2, MC_THREADS = 416;
#else
constexpr int MC_BKV = 128, MC_KST = 2, MC_THREADS = 320;
#endif
constexpr int MC_K_BYTES = 4 * MC_BKV * 128, MC_V_BYTES = MC_BKV * 128;
constexpr int MC_SMEM = MC_KST * MC_K_BYTES + 2 * MC_V_BYTES + 1024 + 4096;
constexpr int MC_P_COL = 64, MC_S_COL = 128, MC_Q_COL = 384;

template <int POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MC_THREADS, 1)
attn_mc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcP p) {
  constexpr int BKV = MC_BKV, KST = MC_KST, D = 256, DV = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sQ = sK + MC_K_BYTES;               // Q staging = K slot 1 (free until the cluster barrier below)
  uint8_t* sV = sK + KST * MC_K_BYTES;
  uint64_t* bars = (uint64_t*)(sV + 2 * MC_V_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2 (one arrival per CTA of the pair)
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2 (one arrival per CTA)
  uint64_t* s_full = bars + 9;        // 2
  uint64_t* p_full = bars + 11;       // 1 (8 softmax warps)
  uint64_t* o_ready = bars + 12;      // 1
  uint64_t* q_tmem = bars + 13;       // 1 (8 softmax warps)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 14);
  float* mxbuf = (float*)(bars + 16);    // [tile parity][half][128 rows]
  float* lbuf = mxbuf + 512;             // [half][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y, b = bh / p.Hh, h = bh - b * p.Hh;
  const int split = blockIdx.z;
  const int t_begin = split * p.tiles_per_split;
  int n = p.ntiles - t_begin;
  if (n > p.tiles_per_split) n = p.tiles_per_split;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&k_full[s], 1);
      tc::mbar_init(&k_empty[s], 2);
      tc::mbar_init(&v_full[s], 1);
      tc::mbar_init(&v_empty[s], 2);
      tc::mbar_init(&s_full[s], 1);
    }
    tc::mbar_init(p_full, 8);
    tc::mbar_init(o_ready, 1);
    tc::mbar_init(q_tmem, 8);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncwarp();
  tc::cluster_arrive();               // both CTAs' barriers exist before anybody multicasts into them
  tc::cluster_wait();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // every tile is loaded as two half-tiles of 64 keys: this CTA fetches half `rank` and multicasts it to both CTAs
    auto load_k = [&](int j) {
      const int ks = j & 1;
      tc::mbar_wait(&k_empty[ks], ((uint32_t)(j >> 1) & 1u) ^ 1u);
      MS2_TRACE(j, 14);
      tc::mbar_arrive_expect_tx(&k_full[ks], MC_K_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tc::tma_load_4d_mc(sK + ks * MC_K_BYTES + c * BKV * 128 + rank * 64 * 128, &tmK, &k_full[ks], c * 64,
                           (t_begin + j) * BKV + (int)rank * 64, h, b, (uint16_t)3);
    };
    auto load_v = [&](int j) {
      const int st = j & 1;
      tc::mbar_wait(&v_empty[st], ((uint32_t)(j >> 1) & 1u) ^ 1u);
      MS2_TRACE(j, 15);
      tc::mbar_arrive_expect_tx(&v_full[st], MC_V_BYTES);
      tc::tma_load_4d_mc(sV + st * MC_V_BYTES + rank * 64 * 128, &tmV, &v_full[st], 0, (t_begin + j) * BKV + (int)rank * 64,
                         h, b, (uint16_t)3);
    };
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(q_full, 4 * BQ * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) tc::tma_load_4d(sQ + c * BQ * 128, &tmQ, q_full, c * 64, q0, h, b);
      load_k(0);
    }
    __syncwarp();
    tc::cluster_arrive();             // Q of BOTH CTAs has left K slot 1 (see the softmax warps)
    tc::cluster_wait();
    if (tc::elect_one()) {
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load_k(j + 1);
        load_v(j);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = tc::make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, DV, 0, 1);
    const uint32_t aK = tc::smem_u32(sK), aV = tc::smem_u32(sV);
    const uint32_t tO = tmem_base, tP = tmem_base + MC_P_COL, tS = tmem_base + MC_S_COL, tQ = tmem_base + MC_Q_COL;
    auto issue_qk = [&](int j) {
      const int st = j & 1, ks = j & 1;
      // S[st] was last read by the softmax of tile j-2 (observed through p_full before P V(j-2) was issued); P has its
      // own columns, so this MMA does not wait for any P V to finish
      MS2_TRACE(j, 0);
      tc::mbar_wait(&k_full[ks], (uint32_t)(j >> 1) & 1u);
      tc::tc_fence_after();
      MS2_TRACE(j, 1);
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t db = tc::desc_kmajor_sw128(aK + ks * MC_K_BYTES + (kk >> 2) * BKV * 128 + (kk & 3) * 32);
          tc::umma_bf16_ts(tS + st * BKV, tQ + kk * 8, db, idesc_qk, kk ? 1u : 0u);
#ifdef MS2_ATTN_TRACE_ISSUE
          if (kk == 3) MS2_TRACE(j, 9);
          if (kk == 7) MS2_TRACE(j, 10);
          if (kk == 11) MS2_TRACE(j, 11);
          if (kk == 15) MS2_TRACE(j, 12);
#endif
        }
        // local signal first: commits retire in order, and the multicast one takes several hundred cycles
        tc::umma_commit(&s_full[st]);
#ifdef MS2_ATTN_TRACE_ISSUE
        MS2_TRACE(j, 13);
#endif
        // the slot is refilled by BOTH CTAs' producers: tell both (only if somebody will wait for it)
        if (j + KST < n) tc::umma_commit_mc(&k_empty[ks], (uint16_t)3);
      }
      __syncwarp();
    };
    tc::cluster_arrive();
    tc::cluster_wait();
    tc::mbar_wait(q_tmem, 0);
    tc::tc_fence_after();
    issue_qk(0);
    if (n > 1) issue_qk(1);
    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      tc::mbar_wait(&v_full[st], (uint32_t)(j >> 1) & 1u);
      MS2_TRACE(j, 2);
      tc::mbar_wait(p_full, (uint32_t)j & 1u);
      tc::tc_fence_after();
      MS2_TRACE(j, 3);
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t db = tc::desc_mnmajor_sw128(aV + st * MC_V_BYTES + kk * 2048, BKV * 128);
          tc::umma_bf16_ts(tO, tP + kk * 8, db, idesc_pv, (j | kk) ? 1u : 0u);
        }
        tc::umma_commit(o_ready);
        if (j + 2 < n) tc::umma_commit_mc(&v_empty[st], (uint16_t)3);
      }
      __syncwarp();
      if (j + 2 < n) issue_qk(j + 2);
    }
#ifdef MS2_ATTN_TRACE_SPIN
  } else if (warp >= 10) {
    // trace builds only: idle observers that time the completion of every S tile / P tile / P V independently of the
    // wake-up latency of the warps that consume them
    tc::cluster_arrive();
    tc::cluster_wait();
    for (int j = 0; j < n; ++j) {
      if (warp == 10) { tc::mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1u); MS2_TRACE(j, 9); }
      if (warp == 11) { tc::mbar_wait(p_full, (uint32_t)j & 1u); MS2_TRACE(j, 10); }
      if (warp == 12) { tc::mbar_wait(o_ready, (uint32_t)j & 1u); MS2_TRACE(j, 11); }
    }
#endif
  } else {
    // ===================== softmax / correction / epilogue (warps 2..9) =====================
    constexpr int HC = BKV / 2;                     // S columns per warp
    const int qtr = warp & 3, half = (warp - 2) >> 2;
    const int row = qtr * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qtr * 32) << 16;
    const uint32_t tO = tmem_base + lane_addr, tS = tmem_base + lane_addr + MC_S_COL;
    float m_used = 0.f, l = 0.f;
    const int last_valid = p.Lk - (p.ntiles - 1) * BKV;

    {  // Q: swizzled staging tile (K slot 1) -> tensor memory; the two warps of a quarter take alternate blocks
      tc::mbar_wait(q_full, 0);
      const uint32_t aQ = tc::smem_u32(sQ);
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {
        if ((blk & 1) != half) continue;
        const int ch = blk >> 1, g0 = (blk & 1) * 4;
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 v = tc::lds128(aQ + ch * BQ * 128 + row * 128 + (((g0 + g) ^ (row & 7)) << 4));
          w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
        }
        tc::tmem_st16(tmem_base + lane_addr + MC_Q_COL + blk * 16, w);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_tmem);
      tc::cluster_arrive();           // K slot 1 may now be overwritten — by either CTA's multicast
      tc::cluster_wait();
    }

    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      if (qtr == 2) MS2_TRACE_SM(j, 4 + half * 5);
      tc::mbar_wait(&s_full[st], (uint32_t)(j >> 1) & 1u);
      tc::tc_fence_after();
      if (qtr == 2) MS2_TRACE_SM(j, 5 + half * 5);
      uint32_t r[HC / 32][32];
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) tc::tmem_ld32(tS + st * BKV + half * HC + c * 32, r[c]);
      tc::tmem_ld_wait();
      if ((t_begin + j == p.ntiles - 1) && (last_valid < BKV)) {
#pragma unroll
        for (int c = 0; c < HC / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (half * HC + c * 32 + i >= last_valid) r[c][i] = 0xff800000u;
      }
      // tile maximum of this warp's 64 columns (two chains: FMNMX3 takes three operands), published to the partner warp
      float mx;
      float* mslot = mxbuf + (j & 1) * 256;
      auto tile_max = [&]() {
        float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mxa = fmaxf(mxa, __uint_as_float(r[0][i]));
          mxb = fmaxf(mxb, __uint_as_float(r[1][i]));
        }
        mx = fmaxf(mxa, mxb) * p.c;
        mslot[half * 128 + row] = mx;
      };
      uint32_t pk[HC / 2];
      float lt;                                       // row sum of this tile's P (this warp's columns)
      // p = 2^(s*c - m): scale-subtract and the row sum on packed fp32 pairs (FFMA2 / FADD2); MUFU.EX2 bounds this loop
      // (tools/ubench/mufu.cu: 1100 clk per 128-key tile per SM sub-partition)
      auto exp_pass = [&](float nm) {
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < HC / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            ffma2(x0, x1, __uint_as_float(r[c][i]), __uint_as_float(r[c][i + 1]), p.c, nm);
            const float p0 = ex2(x0), p1 = ex2(x1);
            fadd2(l0, l1, p0, p1);
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            pk[(c * 32 + i) >> 1] = *(uint32_t*)&hh;
          }
        lt = l0 + l1;
      };
      if (j == 0) {
        tile_max();
        tc::named_bar_sync(1 + qtr, 64);
        m_used = fmaxf(mx, mslot[(half ^ 1) * 128 + row]);
        if (qtr == 2) MS2_TRACE_SM(j, 6 + half * 5);
        exp_pass(-m_used);
      } else {
        // SPECULATE that the running maximum does not grow by more than tau in this tile (true for all but the first
        // few tiles): the exponentials start right after the S load instead of waiting for the row maximum and its
        // exchange between the two warps of the lane quarter; the maximum is checked afterwards, and the tile is
        // recomputed with the new maximum in the rare case that it did grow.
        exp_pass(-m_used);
        tile_max();                                   // same basic block: the FMNMX chain fills MUFU-bound issue slots
        tc::named_bar_sync(1 + qtr, 64);
        mx = fmaxf(mx, mslot[(half ^ 1) * 128 + row]);
        if (qtr == 2) MS2_TRACE_SM(j, 6 + half * 5);
        const bool need = mx > m_used + p.tau;
        if (__any_sync(0xffffffffu, need)) {
          tc::mbar_wait(o_ready, (uint32_t)(j - 1) & 1u);
          tc::tc_fence_after();
          float alpha = 1.f;
          if (need) {
            alpha = ex2(m_used - mx);
            m_used = mx;
            l *= alpha;
          }
          uint32_t o[32];                              // DV = 64: two 32-column blocks, one warp of the quarter each
          tc::tmem_ld32(tO + half * 32, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tc::tmem_st32(tO + half * 32, o);
          tc::tmem_st_wait();
          exp_pass(-m_used);
        }
      }
      l += lt;
      if (qtr == 2) MS2_TRACE_SM(j, 7 + half * 5);
      if (j > 0) tc::mbar_wait(o_ready, (uint32_t)(j - 1) & 1u);      // observe every phase (see attn_tc_kernel)
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) tc::tmem_st16(tO + MC_P_COL + half * (HC / 2) + c * 16, &pk[c * 16]);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
      if (qtr == 2) MS2_TRACE_SM(j, 8 + half * 5);
    }

    lbuf[half * 128 + row] = l;
    tc::named_bar_sync(1 + qtr, 64);
    l += lbuf[(half ^ 1) * 128 + row];
    tc::mbar_wait(o_ready, (uint32_t)(n - 1) & 1u);
    tc::tc_fence_after();
    const int qi = q0 + row;
    uint32_t o[32];
    tc::tmem_ld32(tO + half * 32, o);                // this warp's 32 of the 64 output columns
    tc::tmem_ld_wait();
    if (p.nsplit == 1 && !p.force_part) {
      const float inv = 1.f / l;
      bf16* orow = (bf16*)p.o + (long)b * p.o_bs + (long)h * p.o_hs + (long)qi * p.o_ts;
      if (qi < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          uint32_t* vv = (uint32_t*)&v;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv,
                                                      __uint_as_float(o[g * 8 + 2 * i + 1]) * inv);
            vv[i] = *(uint32_t*)&hh;
          }
          *(uint4*)(orow + half * 32 + g * 8) = v;
        }
      }
    } else {
      const long rix = ((long)split * gridDim.y + bh) * p.Lq + qi;
      float* orow = p.opart + rix * DV;
      if (qi < p.Lq) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *(float4*)(orow + half * 32 + g * 4) =
              make_float4(__uint_as_float(o[g * 4]), __uint_as_float(o[g * 4 + 1]), __uint_as_float(o[g * 4 + 2]),
                          __uint_as_float(o[g * 4 + 3]));
        if (half == 0) *(float2*)(p.ml + rix * 2) = make_float2(m_used, l);
      }
    }
    tc::tc_fence_before();
  }
  __syncwarp();
  tc::cluster_arrive();               // neither CTA leaves while the other may still signal its barriers
  tc::cluster_wait();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// merge the split-KV partials: O = sum_s 2^(m_s - m*) O_s / sum_s 2^(m_s - m*) l_s
template <int D>
__global__ void __launch_bounds__(128)
attn_combine_kernel(const float* __restrict__ opart, const float* __restrict__ ml, bf16* __restrict__ o, long o_bs,
                    long o_hs, long o_ts, int Hh, int Lq, int nsplit, long rows) {
  MS2_PDL_WAIT();
  constexpr int TPR = D / 8;                  // threads per row, 8 columns each
  constexpr int RPB = 128 / TPR > 0 ? 128 / TPR : 1;
  const long rix = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int cg = threadIdx.x % TPR;
  if (threadIdx.x >= RPB * TPR || rix >= rows) return;
  float mstar = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mstar = fmaxf(mstar, ml[((long)s * rows + rix) * 2]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float lsum = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float2 v = *(const float2*)(ml + ((long)s * rows + rix) * 2);
    const float w = exp2f(v.x - mstar);
    lsum += w * v.y;
    const float4* src = (const float4*)(opart + ((long)s * rows + rix) * D + cg * 8);
    const float4 a = src[0], bq = src[1];
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
  }
  const float inv = 1.f / lsum;
  const long bh = rix / Lq;
  const int qi = (int)(rix - bh * Lq);
  const int b = (int)(bh / Hh), h = (int)(bh - (long)b * Hh);
  bf16* dst = o + (long)b * o_bs + (long)h * o_hs + (long)qi * o_ts + cg * 8;
  uint4 v;
  uint32_t* vv = (uint32_t*)&v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
    vv[i] = *(uint32_t*)&hh;
  }
  *(uint4*)dst = v;
}

// ---- split-KV across GPUs: every rank attends over its own share of the memory bank and
//      hands out ONE un-normalised partial per query row; the partials of all ranks are merged exactly as the
//      split-KV partials of one GPU are.
// local splits -> one partial:  m* = max_s m_s,  O* = sum_s 2^(m_s - m*) O_s,  l* = sum_s 2^(m_s - m*) l_s
template <int D>
__global__ void __launch_bounds__(128)
attn_reduce_partials_kernel(const float* __restrict__ opart, const float* __restrict__ ml, float* __restrict__ out_o,
                            float* __restrict__ out_ml, int nsplit, long rows) {
  MS2_PDL_WAIT();
  constexpr int TPR = D / 8, RPB = 128 / TPR;
  const long rix = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int cg = threadIdx.x % TPR;
  if (rix >= rows) return;
  float mstar = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mstar = fmaxf(mstar, ml[((long)s * rows + rix) * 2]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float lsum = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float2 v = *(const float2*)(ml + ((long)s * rows + rix) * 2);
    const float w = exp2f(v.x - mstar);
    lsum += w * v.y;
    const float4* src = (const float4*)(opart + ((long)s * rows + rix) * D + cg * 8);
    const float4 a = src[0], bq = src[1];
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
  }
  float4* dst = (float4*)(out_o + rix * D + cg * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  if (cg == 0) *(float2*)(out_ml + rix * 2) = make_float2(mstar, lsum);
}

// partials of `nparts` ranks (part r at parts_o + r*part_stride / parts_ml + r*part_stride; an empty share is
// (0, -inf, 0)) -> softmax-normalised bf16 rows
template <int D>
__global__ void __launch_bounds__(128)
attn_merge_kernel(const float* __restrict__ parts_o, const float* __restrict__ parts_ml, long part_stride,
                  bf16* __restrict__ o, long o_bs, long o_ts, int Lq, int nparts, long rows) {
  MS2_PDL_WAIT();
  constexpr int TPR = D / 8, RPB = 128 / TPR;
  const long rix = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int cg = threadIdx.x % TPR;
  if (rix >= rows) return;
  float mstar = -INFINITY;
  for (int r = 0; r < nparts; ++r) mstar = fmaxf(mstar, parts_ml[(long)r * part_stride + rix * 2]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float lsum = 0.f;
  for (int r = 0; r < nparts; ++r) {
    const float2 v = *(const float2*)(parts_ml + (long)r * part_stride + rix * 2);
    if (v.y <= 0.f) continue;                                  // empty share
    const float w = exp2f(v.x - mstar);
    lsum += w * v.y;
    const float4* src = (const float4*)(parts_o + (long)r * part_stride + rix * D + cg * 8);
    const float4 a = src[0], bq = src[1];
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
  }
  const float inv = 1.f / lsum;
  const long b = rix / Lq;
  const int qi = (int)(rix - b * Lq);
  uint4 v;
  uint32_t* vv = (uint32_t*)&v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
    vv[i] = *(uint32_t*)&hh;
  }
  *(uint4*)(o + b * o_bs + (long)qi * o_ts + cg * 8) = v;
}

// ---- split-KV across GPUs without a collective call: the kernel that folds this rank's local splits into ONE partial
//      writes it straight into EVERY rank's gather buffer (peer memory over NVLink: plain stores to addresses of the
//      peers' symmetric allocation; slot = this rank) and, when the last CTA has fenced its stores, raises this rank's
//      flag on every peer (st.release.sys).  The consumer is attn_merge_wait_kernel on each rank: it spins on the flags
//      of all ranks (ld.acquire.sys) and merges.  One exchange = (world-1) x 1.03 MiB of NVLink writes per rank.
struct PushP {
  float* dst[8];          // per destination rank: base of MY slot in that rank's gather buffer ([rows*DV] O then [rows*2] (m,l))
  unsigned* flag[8];      // per destination rank: MY flag word there
  unsigned* counter;      // local, zero between launches: CTAs that have fenced their stores
  int world;
  unsigned step;
};

template <int D>
__global__ void __launch_bounds__(128)
attn_reduce_push_kernel(const float* __restrict__ opart, const float* __restrict__ ml, int nsplit, long rows, const PushP pp) {
  MS2_PDL_WAIT();
  constexpr int TPR = D / 8, RPB = 128 / TPR;
  const long rix = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int cg = threadIdx.x % TPR;
  if (rix < rows) {
    float mstar = -INFINITY;
    for (int s = 0; s < nsplit; ++s) mstar = fmaxf(mstar, ml[((long)s * rows + rix) * 2]);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float lsum = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float2 v = *(const float2*)(ml + ((long)s * rows + rix) * 2);
      const float w = exp2f(v.x - mstar);
      lsum += w * v.y;
      const float4* src = (const float4*)(opart + ((long)s * rows + rix) * D + cg * 8);
      const float4 a = src[0], bq = src[1];
      acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
      acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
    }
    const float4 o0 = make_float4(acc[0], acc[1], acc[2], acc[3]), o1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
    for (int r = 0; r < pp.world; ++r) {
      float4* dst = (float4*)(pp.dst[r] + rix * D + cg * 8);
      dst[0] = o0;
      dst[1] = o1;
      if (cg == 0) *(float2*)(pp.dst[r] + rows * D + rix * 2) = make_float2(mstar, lsum);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(pp.counter, 1u);
    if (prev == gridDim.x - 1) {                    // every CTA's stores are fenced: publish
      *pp.counter = 0u;
      __threadfence_system();
      for (int r = 0; r < pp.world; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flag[r]), "r"(pp.step) : "memory");
    }
  }
}

// merge of the `world` partials of one exchange step once every rank's flag shows that step (or a later one)
template <int D>
__global__ void __launch_bounds__(128)
attn_merge_wait_kernel(const float* __restrict__ parts, long part_stride, const unsigned* __restrict__ flags, unsigned step,
                       int world, bf16* __restrict__ o, long o_bs, long o_ts, int Lq, long rows) {
  MS2_PDL_WAIT();
  if (threadIdx.x == 0) {
    const uint64_t t0 = tc::globaltimer_ns();
    for (int r = 0; r < world; ++r) {
      unsigned v;
      uint32_t spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
        if ((int)(v - step) >= 0) break;
        if ((++spins & 0xfffu) == 0 && tc::globaltimer_ns() - t0 > 20000000000ull) {
          printf("medsam2_b200: partial of rank %d (step %u, have %u) never arrived\n", r, step, v);
          __trap();
        }
      } while (true);
    }
  }
  __syncthreads();
  constexpr int TPR = D / 8, RPB = 128 / TPR;
  const long rix = (long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int cg = threadIdx.x % TPR;
  if (rix >= rows) return;
  // peers wrote these lines: read them past the (non-coherent) L1
  float mstar = -INFINITY;
  for (int r = 0; r < world; ++r) mstar = fmaxf(mstar, __ldcg(parts + (long)r * part_stride + rows * D + rix * 2));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float lsum = 0.f;
  for (int r = 0; r < world; ++r) {
    const float2 v = __ldcg((const float2*)(parts + (long)r * part_stride + rows * D + rix * 2));
    if (v.y <= 0.f) continue;                                  // empty share
    const float w = exp2f(v.x - mstar);
    lsum += w * v.y;
    const float4* src = (const float4*)(parts + (long)r * part_stride + rix * D + cg * 8);
    const float4 a = __ldcg(src), bq = __ldcg(src + 1);
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
  }
  const float inv = 1.f / lsum;
  const long b = rix / Lq;
  const int qi = (int)(rix - b * Lq);
  uint4 v;
  uint32_t* vv = (uint32_t*)&v;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
    vv[i] = *(uint32_t*)&hh;
  }
  *(uint4*)(o + b * o_bs + (long)qi * o_ts + cg * 8) = v;
}

int pick_nsplit(int qtiles_total, int ntiles, long ws_rows_bytes_per_split, long ws_bytes) {
  const int sms = tc::sm_count();
  if (qtiles_total >= 2 * sms || ntiles < 4) return 1;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 64 && s <= ntiles / 2; ++s) {
    const int tps = (ntiles + s - 1) / s;
    const int se = (ntiles + tps - 1) / tps;
    if (se != s) continue;
    if (s > 1 && (long)s * ws_rows_bytes_per_split > ws_bytes) break;
    const long ctas = (long)qtiles_total * s;
    const long waves = (ctas + sms - 1) / sms;
    // time ~ waves * (tiles per CTA + fixed prologue/epilogue in tile units) (+ combine traffic)
    const double cost = waves * (tps + 3.0) + (s > 1 ? 0.25 * s : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

template <int D, int DV, int BKV, int KST>
int launch(const void* q, const void* k, const void* v, void* o, long q_bs, long q_hs, long q_ts, long k_bs, long k_hs,
           long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts, int B, int Hh, int Lq, int Lk,
           float scale, void* ws, long ws_bytes, cudaStream_t st, float* part_o = nullptr, float* part_ml = nullptr,
           const PushP* push = nullptr) {
  using C = Cfg<D, DV, BKV, KST>;
  // `push`: the folded partial goes to every rank's gather buffer (attn_reduce_push_kernel) instead of part_o: the
  // attention kernel then always leaves its split partials in the workspace (also when there is a single split)
  float* const ws_ml1 = ws ? (float*)ws + (long)B * Hh * Lq * DV : nullptr;
  if (push) { part_o = (float*)ws; part_ml = ws_ml1; }
  static_assert(C::SMEM <= 227 * 1024, "attention tile does not fit shared memory");
  static_assert(KST >= 2 && KST <= 4, "K ring depth");
  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* m, const void* base, long bs, long hs, long ts, int L, int rows, int width) {
    const uint64_t dims[4] = {(uint64_t)width, (uint64_t)L, (uint64_t)Hh, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ts, (uint64_t)(Hh > 1 ? hs : ts * (long)L), (uint64_t)(B > 1 ? bs : ts * (long)L * Hh)};
    const uint32_t box[4] = {64, (uint32_t)rows, 1, 1};
    return tc::make_tmap_bf16(m, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = mk(&tmQ, q, q_bs, q_hs, q_ts, Lq, BQ, D))) return rc;
  if ((rc = mk(&tmK, k, k_bs, k_hs, k_ts, Lk, BKV, D))) return rc;
  if ((rc = mk(&tmV, v, v_bs, v_hs, v_ts, Lk, BKV, DV))) return rc;

  AttnTcP p;
  p.o = o; p.o_bs = o_bs; p.o_hs = o_hs; p.o_ts = o_ts;
  p.B = B; p.Hh = Hh; p.Lq = Lq; p.Lk = Lk;
  p.c = scale * 1.4426950408889634f;
  {
    static const float tau = []() { const char* e = getenv("MS2_LAZY_TAU"); return e ? (float)atof(e) : LAZY_TAU; }();
    p.tau = tau;
  }
  p.ntiles = (Lk + BKV - 1) / BKV;
  const int qtiles = (Lq + BQ - 1) / BQ;
  const long rows = (long)B * Hh * Lq;
  const long per_split = rows * (DV + 2) * 4;
  p.nsplit = ws ? pick_nsplit(qtiles * B * Hh, p.ntiles, per_split, ws_bytes) : 1;
  p.tiles_per_split = (p.ntiles + p.nsplit - 1) / p.nsplit;
  p.opart = (float*)ws;
  p.ml = p.nsplit > 1 ? (float*)ws + (long)p.nsplit * rows * DV : nullptr;
  p.force_part = part_o ? 1 : 0;

  static const bool two_tiles_ok = []() { const char* e = getenv("MS2_ATTN_TWO_TILES"); return !e || atoi(e) != 0; }();
  static const bool mc_ok = []() { const char* e = getenv("MS2_ATTN_MC"); return e ? atoi(e) != 0 : MC_DEFAULT_ON; }();
  if (D == 256 && DV == 64 && qtiles % 2 == 0 && mc_ok) {
    // CTA pairs sharing K/V through TMA multicast, 128-key tiles (attn_mc_kernel)
    CUtensorMap tmK2, tmV2;
    auto mk2 = [&](CUtensorMap* m, const void* base, long bs, long hs, long ts, int L, int rows, int width) {
      const uint64_t dims[4] = {(uint64_t)width, (uint64_t)L, (uint64_t)Hh, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)ts, (uint64_t)(Hh > 1 ? hs : ts * (long)L), (uint64_t)(B > 1 ? bs : ts * (long)L * Hh)};
      const uint32_t box[4] = {64, (uint32_t)rows, 1, 1};
      return tc::make_tmap_bf16(m, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    if ((rc = mk2(&tmK2, k, k_bs, k_hs, k_ts, Lk, 64, D))) return rc;
    if ((rc = mk2(&tmV2, v, v_bs, v_hs, v_ts, Lk, 64, DV))) return rc;
    p.ntiles = (Lk + MC_BKV - 1) / MC_BKV;
    p.nsplit = ws ? pick_nsplit(qtiles * B * Hh, p.ntiles, per_split, ws_bytes) : 1;
    p.tiles_per_split = (p.ntiles + p.nsplit - 1) / p.nsplit;
    p.ml = p.nsplit > 1 ? (float*)ws + (long)p.nsplit * rows * DV : nullptr;
    if (part_o && p.nsplit == 1) { p.opart = part_o; p.ml = part_ml; }
    // (polynomial share of the exponentials, measured at Lk = 209120: 433 / 443 / 460 / 495 us for 0, 1/8, 1/4, 3/8)
    auto kmc = attn_mc_kernel<0>;
    static bool attr_mc = false;
    if (!attr_mc) {
      MS2_CUDA(cudaFuncSetAttribute(kmc, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_SMEM), "attn_mc attr");
      attr_mc = true;
    }
    dim3 gridm(qtiles, B * Hh, p.nsplit);
    ms2_launch(kmc, gridm, MC_THREADS, MC_SMEM, st, tmQ, tmK2, tmV2, p);
    MS2_CHECK_LAUNCH("attn_mc_kernel");
  } else if (D == 256 && DV == 64 && Lq >= 2 * BQ && two_tiles_ok) {
    // two query tiles per CTA share every K/V tile (half the L2->SM operand traffic per FLOP)
    const int qpairs = (Lq + 2 * BQ - 1) / (2 * BQ);
    p.nsplit = ws ? pick_nsplit(qpairs * B * Hh, p.ntiles, per_split, ws_bytes) : 1;
    p.tiles_per_split = (p.ntiles + p.nsplit - 1) / p.nsplit;
    p.ml = p.nsplit > 1 ? (float*)ws + (long)p.nsplit * rows * DV : nullptr;
    if (part_o && p.nsplit == 1) { p.opart = part_o; p.ml = part_ml; }
    // (softmax variants <1..7>: independent partial maxima / sums and a polynomial share of the exponentials were
    //  measured at Lk = 209120: 525 us for <0> and <1>, 532 / 556 / 575 us with 1/4, 3/8, 1/2 of the exponentials on the
    //  FMA pipe — one softmax warp's instruction stream, not MUFU throughput, is this kernel's critical path)
    static bool attr2 = false;
    if (!attr2) {
      MS2_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM), "attn_tc2 attr");
      attr2 = true;
    }
    dim3 grid2(qpairs, B * Hh, p.nsplit);
    ms2_launch(attn_tc2_kernel<0>, grid2, T2_THREADS, T2_SMEM, st, tmQ, tmK, tmV, p);
    MS2_CHECK_LAUNCH("attn_tc2_kernel");
  } else {
    if (part_o && p.nsplit == 1) { p.opart = part_o; p.ml = part_ml; }
    static const int poly = []() { const char* e = getenv("MS2_ATTN_POLY"); return e ? atoi(e) : ATTN_D96_POLY; }();
    auto kern = attn_tc_kernel<D, DV, BKV, KST, 0>;
    if constexpr (D == 96) {
      switch (poly) {
        case 1: kern = attn_tc_kernel<D, DV, BKV, KST, 1>; break;
        case 2: kern = attn_tc_kernel<D, DV, BKV, KST, 2>; break;
        case 3: kern = attn_tc_kernel<D, DV, BKV, KST, 3>; break;
        case 4: kern = attn_tc_kernel<D, DV, BKV, KST, 4>; break;
        default: break;
      }
    }
    static decltype(kern) attr_set = nullptr;
    if (attr_set != kern) {
      MS2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM), "attn_tc attr");
      attr_set = kern;
    }
    dim3 grid(qtiles, B * Hh, p.nsplit);
    ms2_launch(kern, grid, NUM_THREADS, C::SMEM, st, tmQ, tmK, tmV, p);
    MS2_CHECK_LAUNCH("attn_tc_kernel");
  }
  if (push) {
    constexpr int TPR = DV / 8, RPB = 128 / TPR;
    ms2_launch(attn_reduce_push_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, (const float*)p.opart, (const float*)p.ml,
               p.nsplit, rows, *push);
    MS2_CHECK_LAUNCH("attn_reduce_push_kernel");
  } else if (part_o) {
    if (p.nsplit > 1) {
      constexpr int TPR = DV / 8, RPB = 128 / TPR;
      ms2_launch(attn_reduce_partials_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, p.opart, p.ml, part_o, part_ml, p.nsplit, rows);
      MS2_CHECK_LAUNCH("attn_reduce_partials_kernel");
    }
  } else if (p.nsplit > 1) {
    constexpr int TPR = DV / 8, RPB = 128 / TPR;
    ms2_launch(attn_combine_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, p.opart, p.ml, (bf16*)o, o_bs, o_hs, o_ts, Hh, Lq,
                                                                 p.nsplit, rows);
    MS2_CHECK_LAUNCH("attn_combine_kernel");
  }
  return MS2_OK;
}

}  // namespace

bool ms2_attention_tc_supported(int dt, long q_hs, long q_ts, long k_hs, long k_ts, long v_hs, long v_ts, long o_hs,
                                long o_ts, int Hh, int Lq, int Lk, int D, int DV) {
  if (dt != MS2_BF16) return false;
  if (!(((D == 64 || D == 96 || D == 128 || D == 256) && DV == D) || (D == 256 && DV == 64))) return false;
  if (Lq < 64 || Lk < 64) return false;
  if ((q_ts | k_ts | v_ts | o_ts) % 8) return false;
  if (Hh > 1 && ((q_hs | k_hs | v_hs | o_hs) % 8)) return false;
  return true;
}

int ms2_attention_tc_launch(const void* q, const void* k, const void* v, void* o, long q_bs, long q_hs, long q_ts,
                            long k_bs, long k_hs, long k_ts, long v_bs, long v_hs, long v_ts, long o_bs, long o_hs,
                            long o_ts, int B, int Hh, int Lq, int Lk, int D, int DV, float scale, void* ws,
                            long ws_bytes, cudaStream_t st) {
  MS2_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                    ((uintptr_t)o % 16 == 0) && (!ws || (uintptr_t)ws % 16 == 0),
                "attention_tc: pointers must be 16-byte aligned");
  MS2_CHECK_ARG(B == 1 || ((q_bs | k_bs | v_bs | o_bs) % 8 == 0), "attention_tc: batch strides must be multiples of 8");
#define MS2_ATTN_ARGS q, k, v, o, q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts, B, Hh, Lq, Lk, scale, ws, ws_bytes, st
  if (D == 256 && DV == 64) return launch<256, 64, 64, 4>(MS2_ATTN_ARGS);
  if (DV == D) {
    switch (D) {
      case 64: return launch<64, 64, 128, 2>(MS2_ATTN_ARGS);
      case 96: return launch<96, 96, 128, 2>(MS2_ATTN_ARGS);
      case 128: return launch<128, 128, 128, 2>(MS2_ATTN_ARGS);
      case 256: return launch<256, 256, 64, 2>(MS2_ATTN_ARGS);
    }
  }
#undef MS2_ATTN_ARGS
  ms2_set_error("attention_tc: unsupported head dims %d/%d", D, DV);
  return MS2_ERR_UNSUPPORTED;
}

// ---- split-KV across GPUs: partial attention over this rank's keys / merge of all ranks' partials (D=256, DV=64)
int ms2_attention_dv_partial_launch(const void* q, const void* k, const void* v, float* part_o, float* part_ml, long q_bs,
                                    long q_ts, long k_bs, long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk,
                                    float scale, void* ws, long ws_bytes, cudaStream_t st) {
  MS2_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                    ((uintptr_t)part_o % 16 == 0) && ((uintptr_t)part_ml % 8 == 0) && (!ws || (uintptr_t)ws % 16 == 0),
                "attention_dv_partial: pointers must be 16-byte aligned");
  return launch<256, 64, 64, 4>(q, k, v, nullptr, q_bs, 256, q_ts, k_bs, 256, k_ts, v_bs, 64, v_ts, 0, 0, 0, B, 1, Lq, Lk,
                                scale, ws, ws_bytes, st, part_o, part_ml);
}

int ms2_attention_merge_launch(const float* parts_o, const float* parts_ml, long part_stride, void* o, long o_bs, long o_ts,
                               int B, int Lq, int nparts, cudaStream_t st) {
  MS2_CHECK_ARG(((uintptr_t)parts_o % 16 == 0) && ((uintptr_t)parts_ml % 8 == 0) && ((uintptr_t)o % 16 == 0) &&
                    part_stride % 4 == 0 && o_ts % 8 == 0 && (B == 1 || o_bs % 8 == 0),
                "attention_merge: alignment");
  const long rows = (long)B * Lq;
  constexpr int DV = 64, TPR = DV / 8, RPB = 128 / TPR;
  ms2_launch(attn_merge_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, parts_o, parts_ml, part_stride, (bf16*)o, o_bs, o_ts, Lq,
                                                             nparts, rows);
  MS2_CHECK_LAUNCH("attn_merge_kernel");
  return MS2_OK;
}

int ms2_attention_dv_partial_push_launch(const void* q, const void* k, const void* v, long q_bs, long q_ts, long k_bs,
                                         long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk, float scale, void* ws,
                                         long ws_bytes, const void* const* h_dst, const void* const* h_flag, int world,
                                         unsigned step, void* counter, cudaStream_t st) {
  MS2_CHECK_ARG(world >= 1 && world <= 8 && h_dst && h_flag && counter, "attention_dv_partial_push: 1..8 ranks");
  const long rows = (long)B * Lq;
  MS2_CHECK_ARG(ws && ws_bytes >= rows * (64 + 2) * 4 && (uintptr_t)ws % 16 == 0, "attention_dv_partial_push: workspace too small");
  PushP pp;
  pp.world = world; pp.step = step; pp.counter = (unsigned*)counter;
  for (int r = 0; r < world; ++r) {
    MS2_CHECK_ARG(h_dst[r] && h_flag[r] && (uintptr_t)h_dst[r] % 16 == 0, "attention_dv_partial_push: bad destination %d", r);
    pp.dst[r] = (float*)h_dst[r];
    pp.flag[r] = (unsigned*)h_flag[r];
  }
  if (Lk <= 0) {                                    // no keys on this rank: the empty partial (0, -inf, 0)
    constexpr int DV = 64, TPR = DV / 8, RPB = 128 / TPR;
    ms2_launch(attn_reduce_push_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, (const float*)ws, (const float*)ws, 0, rows, pp);
    MS2_CHECK_LAUNCH("attn_reduce_push_kernel");
    return MS2_OK;
  }
  MS2_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0),
                "attention_dv_partial_push: pointers must be 16-byte aligned");
  return launch<256, 64, 64, 4>(q, k, v, nullptr, q_bs, 256, q_ts, k_bs, 256, k_ts, v_bs, 64, v_ts, 0, 0, 0, B, 1, Lq, Lk,
                                scale, ws, ws_bytes, st, nullptr, nullptr, &pp);
}

int ms2_attention_merge_wait_launch(const float* parts, long part_stride, const void* flags, unsigned step, int world, void* o,
                                    long o_bs, long o_ts, int B, int Lq, cudaStream_t st) {
  MS2_CHECK_ARG(parts && flags && o && world >= 1 && world <= 8 && (uintptr_t)parts % 16 == 0 && part_stride % 4 == 0 &&
                    (uintptr_t)o % 16 == 0 && o_ts % 8 == 0, "attention_merge_wait: bad args");
  const long rows = (long)B * Lq;
  constexpr int DV = 64, TPR = DV / 8, RPB = 128 / TPR;
  ms2_launch(attn_merge_wait_kernel<DV>, ceil_div(rows, RPB), 128, 0, st, parts, part_stride, (const unsigned*)flags, step, world,
             (bf16*)o, o_bs, o_ts, Lq, rows);
  MS2_CHECK_LAUNCH("attn_merge_wait_kernel");
  return MS2_OK;
}
