// tcgen05 / TMEM / TMA bf16 GEMM with the fused epilogue of ms2_gemm (include/medsam2_b200.h):
//     out[M,N] = residual + colscale * act(A[M,K] @ W[N,K]^T + bias)
// Replaces the cuBLAS addmm / 1x1-conv calls behind every nn.Linear of the hot path (SURVEY §8(a)
// a2,a3,a6,a8,a9).  Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer  : A (128 x 64) and W (BN x 64) boxes, 128B-swizzled, ring of `stages`
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into one of two
//                               TMEM accumulator stages; tcgen05.commit releases smem slots / signals epilogue
//   warps 2..   epilogue      : 8 or 16 warps in groups of 4 (one warp per TMEM lane quarter); groups take
//                               alternate 128-byte column chunks: tcgen05.ld -> bias / act / colscale / residual
//                               -> private 128B-swizzled 4 KB staging buffer -> per-warp TMA store (coalesced,
//                               M/N tails clipped by hardware), overlapping the next tile's main loop
// Both operands are K-major, so no transposes exist anywhere; M/N/K tails are handled by TMA
// out-of-bounds zero fill / clipping.  BN (32..256) is a run-time choice per problem shape.
// Most GEMMs of this model have K <= 384 and are HBM-bound: the epilogue, not the MMA loop, sets their speed.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BK = 64, MAX_STAGES = 8, ACC_STAGES = 2, TMEM_COLS = 512;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int MAX_EPI_WARPS = 16, EPI_BUF_BYTES = 32 * 128;     // per warp: one staging buffer of [32 rows][128 B]
constexpr int MAX_THREADS = 64 + MAX_EPI_WARPS * 32;
constexpr int SMEM_TOTAL = 226 * 1024;

struct GemmP {
  const float* bias;
  const float* colscale;
  const float* residual;
  long ldr;
  int M, N, K, BN, tiles_n, tiles, num_kb, stages, epi_warps, epi_bufs;
  int pair;       // CTA pairs (cluster of 2) share every W tile through TMA multicast; tiles (2*mp + rank, n)
};

template <int ACT, bool F32OUT>
__device__ __forceinline__ float apply_act(float v) {
  if (ACT == 1) return F32OUT ? gelu_erf_fast(v) : gelu_erf_tanh(v);
  if (ACT == 2) return fmaxf(v, 0.f);
  if (ACT == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

// 32 fp32 values (bias or colscale) for columns [n, n+32), zero beyond N (N % 4 == 0); issued as 8 independent
// 16-byte loads at the top of a chunk so their latency hides under the tcgen05.ld
__device__ __forceinline__ void load_cols32(const float* __restrict__ src, int n, int N, float4 (&b)[8]) {
  if (n + 32 <= N) {
#pragma unroll
    for (int g = 0; g < 8; ++g) b[g] = __ldg((const float4*)(src + n + g * 4));
  } else {
#pragma unroll
    for (int g = 0; g < 8; ++g)
      b[g] = (n + g * 4 < N) ? __ldg((const float4*)(src + n + g * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool F32OUT, int ACT, bool RES>
__global__ void __launch_bounds__(MAX_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const GemmP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
  uint8_t* staging = smem + (size_t)p.stages * stage_bytes;               // epi_warps x epi_bufs x 4 KB, 1024-aligned
  uint64_t* full = (uint64_t*)(staging + p.epi_warps * p.epi_bufs * EPI_BUF_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint32_t* tmem_ptr = (uint32_t*)(acc_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // i-th tile of this CTA (-1: none).  Plain: tiles blockIdx.x, +gridDim.x, ...  Pair mode: cluster cl works on the
  // tile pairs cl, cl + #clusters, ...; pair tp = (m-tile pair mp, n-tile) and this CTA takes m-tile 2*mp + rank, so
  // the two CTAs of a cluster always need the SAME W tile at the same time.
  const uint32_t rank = p.pair ? tc::cluster_ctarank() : 0u;
  const int ncl = p.pair ? (int)(gridDim.x >> 1) : 0, cl = (int)(blockIdx.x >> 1);
  const int pair_tiles = p.tiles >> 1;
  auto tile_of = [&](int i) -> int {
    if (!p.pair) {
      const long t = (long)blockIdx.x + (long)i * gridDim.x;
      return t < p.tiles ? (int)t : -1;
    }
    const long tp = (long)cl + (long)i * ncl;
    if (tp >= pair_tiles) return -1;
    const int mp = (int)(tp / p.tiles_n), n = (int)(tp - (long)mp * p.tiles_n);
    return (2 * mp + (int)rank) * p.tiles_n + n;
  };
  int my_tiles = 0;
  while (tile_of(my_tiles) >= 0) ++my_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    tc::prefetch_tmap(&tmO);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], p.pair ? 2 : 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], p.epi_warps);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, TMEM_COLS);
  tc::tc_fence_before();
  if (p.pair) {                       // both CTAs' barriers exist before anybody multicasts into them
    __syncwarp();
    tc::cluster_arrive();
    tc::cluster_wait();
  } else {
    __syncthreads();
  }
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();      // barriers, tensor-memory allocation and descriptor prefetch above overlap the preceding kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int half_rows = p.BN >> 1;
      for (int i = 0, t; (t = tile_of(i)) >= 0; ++i) {
        const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * p.BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          tc::mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
          tc::tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
          if (p.pair)      // this CTA fetches half `rank` of the W tile; the hardware writes it into both CTAs
            tc::tma_load_2d_mc(sa + A_STAGE_BYTES + rank * half_rows * 128, &tmW, &full[stage], kb * BK,
                               n0 + (int)rank * half_rows, (uint16_t)3);
          else
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
    long it = 0;                                     // k-block iterations issued so far
    const long it_total = (long)my_tiles * p.num_kb;
    for (int i = 0; tile_of(i) >= 0; ++i) {
      tc::mbar_wait(&acc_empty[as], aphase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            tc::umma_bf16(d_tmem, tc::desc_kmajor_sw128(sa + k * 32), tc::desc_kmajor_sw128(sb + k * 32), idesc,
                          (kb | k) ? 1u : 0u);
          }
          // pair mode: the slot is refilled by both CTAs' multicasts, so both must hear about it — unless nobody
          // will wait for this slot again (the peer may be gone by the time a trailing signal lands)
          if (!p.pair) tc::umma_commit(&empty[stage]);
          else if (it + p.stages < it_total) tc::umma_commit_mc(&empty[stage], (uint16_t)3);
          if (kb == p.num_kb - 1) tc::umma_commit(&acc_full[as]);
        }
        __syncwarp();
        ++it;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2 .. 2+epi_warps) =====================
    // Every warp owns the 32 accumulator rows of its TMEM lane quarter and a private 4 KB staging buffer, so
    // the only synchronisation on this path is __syncwarp(): no CTA-level barrier, stores are per-warp TMA boxes.
    // With a residual, the warp first copies the residual box into the staging buffer with coalesced 16-byte
    // loads (4 rows x 128 B per instruction), then every lane updates its own row in place.
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2, ngrp = p.epi_warps >> 2;
    constexpr int CPC = F32OUT ? 32 : 64;        // columns per 128-byte chunk
    const int nbuf = p.epi_bufs;                 // staging buffers of this warp (2: stores / residual fetches overlap)
    const uint32_t sbuf0 = tc::smem_u32(staging + (warp - 2) * nbuf * EPI_BUF_BYTES);
    auto tile_nch = [&](int t) {
      int ncols = p.N - (t % p.tiles_n) * p.BN;
      if (ncols > p.BN) ncols = p.BN;
      return (ncols + CPC - 1) / CPC;
    };
    // residual box [32 rows][32 fp32] of chunk c of tile t -> staging buffer, as asynchronous 16-byte global->shared
    // copies (4 rows x 128 B per instruction, coalesced); only fp32 outputs carry a residual (CPC == 32)
    auto fetch_residual = [&](int t, int c, uint32_t sb) {
      const int m0 = (t / p.tiles_n) * BM + q * 32, nb = (t % p.tiles_n) * p.BN + c * CPC;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rrow = i * 4 + (lane >> 3), cc = lane & 7;
        const uint32_t dst = sb + rrow * 128 + ((cc ^ (rrow & 7)) << 4);
        if ((m0 + rrow < p.M) && (nb + cc * 4 < p.N))
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                       "l"(p.residual + (long)(m0 + rrow) * p.ldr + nb + cc * 4) : "memory");
        else
          tc::sts128(dst, make_uint4(0, 0, 0, 0));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the chunk this warp handles after (t, c): same tile, or the first one of a later tile (groups rotate per tile)
    // (ti = index into this CTA's tile sequence, tile_of(ti) = the tile)
    auto next_chunk = [&](int ti, int c, int rot, int& tin, int& cn) -> bool {
      tin = ti; cn = c + ngrp;
      while (cn >= tile_nch(tile_of(tin))) {
        ++tin;
        if (tile_of(tin) < 0) return false;
        rot = (rot + 1 == ngrp) ? 0 : rot + 1;
        cn = rot;
      }
      return true;
    };
    const bool has_res = RES && p.residual;
    int as = 0, it = 0, rot = grp;
    uint32_t aphase = 0;
    bool fetched = false;                          // the residual of the upcoming chunk is already on its way
    if (has_res && nbuf > 1) {
      int c0 = rot - ngrp, tin, cn;                   // "chunk before the first": next_chunk finds the first one
      if (tile_of(0) >= 0 && next_chunk(0, c0, rot, tin, cn)) { fetch_residual(tile_of(tin), cn, sbuf0); fetched = true; }
    }
    for (int ti = 0, t; (t = tile_of(ti)) >= 0; ++ti) {
      const int m0 = (t / p.tiles_n) * BM + q * 32, n0 = (t % p.tiles_n) * p.BN;
      const int nch = tile_nch(t);
      bool waited = false;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      for (int c = rot; c < nch; c += ngrp, ++it) {   // groups rotate over tiles so odd chunk counts balance
        const uint32_t sbuf = sbuf0 + (nbuf > 1 ? (it & 1) * EPI_BUF_BYTES : 0), srow = sbuf + lane * 128;
        if (nbuf > 1) {
          // the OTHER buffer is about to be refilled (residual of the next chunk) / this one was last read by the
          // store two chunks ago: at most the most recent store may still be reading
          if (it >= 1) {
            if (tc::elect_one()) {
              if (has_res) tc::tma_store_wait_read<0>();
              else tc::tma_store_wait_read<1>();
            }
            __syncwarp();
          }
          if (has_res) {
            if (!fetched) fetch_residual(t, c, sbuf);          // (only when the look-ahead found nothing)
            int tin, cn;
            const bool more = next_chunk(ti, c, rot, tin, cn);
            if (more) {
              fetch_residual(tile_of(tin), cn, sbuf0 + ((it + 1) & 1) * EPI_BUF_BYTES);
              asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
              asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            fetched = more;
            __syncwarp();
          }
        } else {
          if (it >= 1) {                              // the previous store of this warp has read the buffer
            if (tc::elect_one()) tc::tma_store_wait_read<0>();
            __syncwarp();
          }
          if (has_res) {
            fetch_residual(t, c, sbuf);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
          }
        }
        if (!waited) {
          tc::mbar_wait(&acc_full[as], aphase);
          tc::tc_fence_after();
          waited = true;
        }
        const int nb = n0 + c * CPC;
#pragma unroll
        for (int hlf = 0; hlf < CPC / 32; ++hlf) {
          const int nh = nb + hlf * 32;
          uint32_t acc[32];
          tc::tmem_ld32(taddr + c * CPC + hlf * 32, acc);
          float4 bb[8];
          if (p.bias) load_cols32(p.bias, nh, p.N, bb);
          tc::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float v[4] = {__uint_as_float(acc[g * 4]), __uint_as_float(acc[g * 4 + 1]), __uint_as_float(acc[g * 4 + 2]),
                          __uint_as_float(acc[g * 4 + 3])};
            if (p.bias) { v[0] += bb[g].x; v[1] += bb[g].y; v[2] += bb[g].z; v[3] += bb[g].w; }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = apply_act<ACT, F32OUT>(v[j]);
            if (RES && p.colscale && nh + g * 4 < p.N) {      // rare (layer-scale of the memory-encoder fuser)
              const float4 cs = __ldg((const float4*)(p.colscale + nh + g * 4));
              v[0] *= cs.x; v[1] *= cs.y; v[2] *= cs.z; v[3] *= cs.w;
            }
            if (F32OUT) {
              const uint32_t slot = srow + ((g ^ (lane & 7)) << 4);
              if (has_res) {
                const float4 r4 = tc::lds128f(slot);
                v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
              }
              tc::sts128f(slot, v[0], v[1], v[2], v[3]);
            } else {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
              acc[g * 2] = *(uint32_t*)&h0;           // repack in place: 8 columns -> 4 words
              acc[g * 2 + 1] = *(uint32_t*)&h1;
            }
          }
          if (!F32OUT) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tc::sts128(srow + (((hlf * 4 + g) ^ (lane & 7)) << 4), acc[g * 4], acc[g * 4 + 1], acc[g * 4 + 2],
                         acc[g * 4 + 3]);
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (m0 < p.M && tc::elect_one()) {
          tc::tma_store_2d(&tmO, (const void*)(staging + (sbuf - tc::smem_u32(staging))), nb, m0);
          tc::tma_store_commit();
        }
      }
      rot = (rot + 1 == ngrp) ? 0 : rot + 1;
      if (!waited) {                                 // no chunk of this tile was mine: still observe the phase
        tc::mbar_wait(&acc_full[as], aphase);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
    if (tc::elect_one()) tc::tma_store_wait<0>();
  }
  tc::tc_fence_before();
  if (p.pair) {                       // neither CTA leaves while the other may still multicast into it
    __syncwarp();
    tc::cluster_arrive();
    tc::cluster_wait();
  } else {
    __syncthreads();
  }
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int pick_bn(int M, int N, int K, int o_dt) {
  const int sms = tc::sm_count();
  const long tm = (M + BM - 1) / BM;
  const long kb = (K + BK - 1) / BK;
  const int step = o_dt == MS2_BF16 ? 64 : 32;          // whole 128-byte output chunks per tile
  int best = step;
  double best_cost = 1e30;
  const int nmax = ((N + step - 1) / step) * step;
  for (int bn = step; bn <= 256 && bn <= nmax; bn += step) {
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
  return a_dt == MS2_BF16 && w_dt == MS2_BF16 && M >= 64 && K % 8 == 0 && K >= 16 && lda % 8 == 0 && N % 4 == 0 &&
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
  p.bias = bias; p.colscale = colscale; p.residual = residual; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  static const int bn_env = []() { const char* e = getenv("MS2_GEMM_BN"); return e ? atoi(e) : 0; }();   // tuning aid
  p.BN = bn_env > 0 ? bn_env : pick_bn(M, N, K, o_dt);
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.tiles = ((M + BM - 1) / BM) * p.tiles_n;
  p.num_kb = (K + BK - 1) / BK;
  const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
  // short-K problems are epilogue/HBM-bound: spend shared memory on 16 epilogue warps instead of pipeline depth
  p.epi_warps = p.num_kb <= 8 ? 16 : 8;
  // two staging buffers per epilogue warp (the TMA store of one chunk and the residual fetch of the next overlap the
  // arithmetic of the current one) as long as >= 3 pipeline stages still fit
  static const int bufs_env = []() { const char* e = getenv("MS2_GEMM_EPI_BUFS"); return e ? atoi(e) : 1; }();
  p.epi_bufs = bufs_env >= 2 ? 2 : 1;
  if (p.epi_bufs == 2 && (SMEM_TOTAL - 1024 - 256 - p.epi_warps * 2 * EPI_BUF_BYTES) / stage_bytes < 3) {
    if (p.epi_warps == 16 && (SMEM_TOTAL - 1024 - 256 - 8 * 2 * EPI_BUF_BYTES) / stage_bytes >= 3) p.epi_warps = 8;
    else p.epi_bufs = 1;
  }
  const int stage_budget = SMEM_TOTAL - 1024 - 256 - p.epi_warps * p.epi_bufs * EPI_BUF_BYTES;
  p.stages = stage_budget / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  const size_t smem = (size_t)p.stages * stage_bytes + p.epi_warps * p.epi_bufs * EPI_BUF_BYTES + 1024 + 256;

  CUtensorMap tmA, tmW, tmO;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)lda};
    const uint32_t box[2] = {BK, BM};
    if ((rc = tc::make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K};
    const uint32_t box[2] = {BK, (uint32_t)p.BN};
    if ((rc = tc::make_tmap_bf16(&tmW, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldo};
    if (o_dt == MS2_F32) {
      const uint32_t box[2] = {32, 32};
      rc = tc::make_tmap_f32(&tmO, out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      const uint32_t box[2] = {64, 32};
      rc = tc::make_tmap_bf16(&tmO, out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc) return rc;
  }
  // CTA pairs: worth it where the W tile dominates the L2 -> SM operand traffic of a tile (wide N tile, K >= 256) and
  // there is more than one wave of tiles; MS2_GEMM_PAIR=0 disables, =2 forces (whenever the shape allows it)
  static const int pair_env = []() { const char* e = getenv("MS2_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  const int tiles_m = (M + BM - 1) / BM;
  const bool pair_possible = tiles_m % 2 == 0 && p.BN % 16 == 0 && p.tiles >= 2;
  p.pair = 0;
  if (pair_possible && pair_env == 2) p.pair = 1;
  else if (pair_possible && pair_env == 1 && p.BN >= 128 && K >= 4096 && p.tiles >= 2 * tc::sm_count()) p.pair = 1;
  CUtensorMap tmWp;
  if (p.pair) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K};
    const uint32_t box[2] = {BK, (uint32_t)(p.BN / 2)};
    if ((rc = tc::make_tmap_bf16(&tmWp, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  int grid = p.tiles < tc::sm_count() ? p.tiles : tc::sm_count();
  if (p.pair) {
    int ncl = tc::sm_count() / 2;
    if (ncl > p.tiles / 2) ncl = p.tiles / 2;
    grid = 2 * ncl;
  }
  const bool res = residual || colscale;
  MS2_CHECK_ARG(!residual || o_dt == MS2_F32, "gemm_tc: a residual needs an fp32 output");
#define MS2_GEMM_TC(F, A, R)                                                                                   \
  do {                                                                                                         \
    auto kern = gemm_tc_kernel<F, A, R>;                                                                       \
    static bool attr_set = false;                                                                              \
    if (!attr_set) {                                                                                           \
      MS2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL), "gemm_tc attr"); \
      attr_set = true;                                                                                         \
    }                                                                                                          \
    ms2_launch_cluster(kern, p.pair ? 2 : 1, grid, 64 + p.epi_warps * 32, smem, st, tmA, p.pair ? tmWp : tmW, tmO, p);  \
  } while (0)
#define MS2_GEMM_TC_ACT(F, R)                                                                                  \
  do {                                                                                                         \
    if (act == 0) MS2_GEMM_TC(F, 0, R);                                                                        \
    else if (act == 1) MS2_GEMM_TC(F, 1, R);                                                                   \
    else if (act == 2) MS2_GEMM_TC(F, 2, R);                                                                   \
    else MS2_GEMM_TC(F, 3, R);                                                                                 \
  } while (0)
  if (o_dt == MS2_F32) {
    if (res) MS2_GEMM_TC_ACT(true, true);
    else MS2_GEMM_TC_ACT(true, false);
  } else {
    if (res) MS2_GEMM_TC_ACT(false, true);
    else MS2_GEMM_TC_ACT(false, false);
  }
#undef MS2_GEMM_TC_ACT
#undef MS2_GEMM_TC
  MS2_CHECK_LAUNCH("gemm_tc_kernel");
  return MS2_OK;
}
