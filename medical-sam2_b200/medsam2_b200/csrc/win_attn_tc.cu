// Hiera windowed attention on tcgen05 (bf16, head dim 96): hieradet.py:58-83,136-159 +
// backbones/utils.py:16-62.  Window partition, zero-pad-AFTER-norm semantics (pad tokens enter softmax
// with k = b_k, v = b_v: SURVEY §0 finding 6), the 2x2 q max-pool and unpartition+crop are folded into the
// gather that builds the swizzled shared-memory operands and into the scatter of the epilogue, so no padded
// or re-ordered copy of the tokens ever exists in HBM.
//
// One CTA = one group of G windows of one head packed into a single M=128 tile (block-diagonal mask):
//   all warps : K, V (NK <= 256 key slots) gathered ONCE -> SW128 smem; per query tile: Q (<=128 rows, optional
//               2x2 max-pool) -> SW128 smem.  Row -> token offsets come from a small table built once per row.
//   1 thread  : S = Q K^T  (tcgen05.mma M=128, N=NK, K=96)                       -> TMEM
//   warps 0-7 : row softmax restricted to the row's own window; two warps per TMEM lane quarter split the
//               column chunks their rows can touch; P (bf16 pairs) overwrites S in tensor memory
//   1 thread  : O = P V    (M=128, N=96, K=NK; P from tensor memory, V MN-major) -> TMEM
//   warps 0-7 : O / l -> bf16 -> scatter to [B,Ho,Wo,heads*D]
// Windows of 196 tokens (stage 3) run their two query tiles in the same CTA, sharing the gathered K/V.  Small-window
// configurations need <= 113 KB of shared memory and 256 TMEM columns, so two CTAs share an SM and overlap phases.
#include "tc_common.cuh"

namespace {

constexpr int WD = 96;            // head dim
constexpr int WDCH = 2;           // 64-column chunks
constexpr int WCH16 = WD / 8;     // 16-byte chunks per row (12)
constexpr int NT = 256;

struct WinP {
  const bf16* qkv;
  const float* bias;
  bf16* out;
  int B, H, W, heads, ws, qpool, nwx, nwy, Ho, Wo, dim_out;
  int lq_w, lk_w, G, q_tiles, NK, groups;
  int tmem_cols, o_col;
  float c;
};

__device__ __forceinline__ float ex2w(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = (const __nv_bfloat162*)&a;
  const __nv_bfloat162* pb = (const __nv_bfloat162*)&b;
  __nv_bfloat162* pr = (__nv_bfloat162*)&r;
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MAXC = column chunks (of 32) a warp half may own: 2 when NK <= 128 (<= 128 registers/thread, two CTAs per SM),
// 4 for the 196-key windows.
template <int MAXC>
__global__ void __launch_bounds__(NT, MAXC == 2 ? 2 : 1)
win_attn_tc_kernel(const WinP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kv_bytes = WDCH * p.NK * 128;
  uint8_t* sQ = smem;                               // q_tiles x 2 x [128][128 B]
  uint8_t* sK = sQ + p.q_tiles * WDCH * 128 * 128;  // 2 x [NK][128 B]
  uint8_t* sV = sK + kv_bytes;
  uint8_t* tail = sV + kv_bytes;
  uint4* sBias = (uint4*)tail;                      // [3][12] 16-byte chunks of the bf16 bias of this head
  int* rowtab = (int*)(tail + 3 * WCH16 * 16);      // [2*128*4 + 256] token element offsets (-1 pad -> bias, -2 none)
  float* mxs = (float*)(rowtab + 2 * 128 * 4 + 256);  // [2 halves][128 rows] row-max / row-sum exchange
  uint64_t* bar_s = (uint64_t*)(mxs + 256);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_ptr = (uint32_t*)(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nwin = p.nwx * p.nwy;
  const int w0 = grp * p.G;
  const long C3 = 3L * p.dim_out;
  const bf16* base = p.qkv + (long)b * p.H * p.W * C3 + h * WD;
  const uint32_t aQ = tc::smem_u32(sQ), aK = tc::smem_u32(sK), aV = tc::smem_u32(sV);

  if (tid == 0) {
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_o, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
  if (tid < 3 * WCH16) {
    const int which = tid / WCH16, ch = tid - which * WCH16;
    const float* bs = p.bias + which * p.dim_out + h * WD + ch * 8;
    uint4 v;
    __nv_bfloat162* pv = (__nv_bfloat162*)&v;
#pragma unroll
    for (int i = 0; i < 4; ++i) pv[i] = __floats2bfloat162_rn(bs[2 * i], bs[2 * i + 1]);
    sBias[tid] = v;
  }
  // K/V row table: key slot -> element offset of its token (the divisions happen once per row, not per chunk)
  for (int r = tid; r < p.NK; r += NT) {
    const int wl = r / p.lk_w, j = r - wl * p.lk_w, w = w0 + wl;
    int off = -2;
    if (wl < p.G && w < nwin) {
      const int wy = w / p.nwx, wx = w - wy * p.nwx;
      const int y = wy * p.ws + j / p.ws, x = wx * p.ws + j % p.ws;
      off = (y < p.H && x < p.W) ? (int)(((long)y * p.W + x) * C3) : -1;
    }
    rowtab[1024 + r] = off;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  MS2_PDL_WAIT();      // barriers, tensor-memory allocation and descriptor prefetch above overlap the preceding kernel
  const uint32_t tS = tmem_base, tO = tmem_base + (uint32_t)p.o_col;

  // Q row table of query tile qt (4 source positions with the 2x2 max-pool, 1 without)
  auto build_q_rowtab = [&](int qt) {
    if (tid < 128) {
      int wl, i;
      if (p.q_tiles > 1) { wl = 0; i = qt * 128 + tid; }
      else { wl = tid / p.lq_w; i = tid - wl * p.lq_w; }
      const int w = w0 + wl;
      int o4[4] = {-2, -2, -2, -2};
      if (wl < p.G && w < nwin && i < p.lq_w) {
        const int wy = w / p.nwx, wx = w - wy * p.nwx;
        if (!p.qpool) {
          const int y = wy * p.ws + i / p.ws, x = wx * p.ws + i % p.ws;
          o4[0] = (y < p.H && x < p.W) ? (int)(((long)y * p.W + x) * C3) : -1;
        } else {
          const int hw = p.ws >> 1;
          const int y = wy * p.ws + 2 * (i / hw), x = wx * p.ws + 2 * (i % hw);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int yy = y + (q >> 1), xx = x + (q & 1);
            o4[q] = (yy < p.H && xx < p.W) ? (int)(((long)yy * p.W + xx) * C3) : -1;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) rowtab[(qt & 1) * 512 + tid * 4 + q] = o4[q];
    }
  };
  // Q tile qt -> sQ[qt]: without pooling every 16-byte piece is an asynchronous global->shared copy (no register
  // staging, all pieces of a thread in flight at once); with the 2x2 max-pool the four source rows go through registers
  auto gather_q = [&](int qt) {
    const int* rt = rowtab + (qt & 1) * 512;
    const uint32_t dstQ = aQ + (uint32_t)qt * (WDCH * 128 * 128);
    if (!p.qpool) {
      for (int idx = tid; idx < 128 * WCH16; idx += NT) {
        const int rr = idx / WCH16, ch = idx - rr * WCH16;
        const uint32_t dst = dstQ + (ch >> 3) * 128 * 128 + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4);
        const int o0 = rt[rr * 4];
        if (o0 >= 0) cp_async16(dst, base + o0 + ch * 8);
        else tc::sts128(dst, o0 == -1 ? sBias[ch] : make_uint4(0, 0, 0, 0));
      }
    } else {
      // valid (pooled) query rows are the first nv rows of the tile; two pieces per thread and pass, so that eight
      // independent 16-byte loads are in flight before the first max
      int nv = 128;
      if (p.q_tiles == 1) {
        int nw = nwin - w0;
        if (nw > p.G) nw = p.G;
        nv = nw * p.lq_w;
      }
      auto dst_of = [&](int rr, int ch) { return dstQ + (ch >> 3) * 128 * 128 + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4); };
      for (int idx = tid + nv * WCH16; idx < 128 * WCH16; idx += NT)          // rows without a query: zeros
        tc::sts128(dst_of(idx / WCH16, idx % WCH16), make_uint4(0, 0, 0, 0));
      for (int i0 = tid; i0 < nv * WCH16; i0 += 2 * NT) {
        uint4 v[2][4];
        int rrs[2], chs[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = i0 + u * NT;
          on[u] = idx < nv * WCH16;
          rrs[u] = on[u] ? idx / WCH16 : 0;
          chs[u] = on[u] ? idx - rrs[u] * WCH16 : 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int off = on[u] ? rt[rrs[u] * 4 + q] : -2;
            v[u][q] = off >= 0 ? __ldg((const uint4*)(base + off + chs[u] * 8)) : (off == -1 ? sBias[chs[u]] : make_uint4(0, 0, 0, 0));
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (on[u]) {
            const bool none = rt[rrs[u] * 4] == -2;
            const uint4 r = max_bf16x8(max_bf16x8(v[u][0], v[u][1]), max_bf16x8(v[u][2], v[u][3]));
            tc::sts128(dst_of(rrs[u], chs[u]), none ? make_uint4(0, 0, 0, 0) : r);
          }
      }
    }
  };

  // ---- gather K, V once per CTA (shared by all query tiles of the window group): asynchronous 16-byte copies, so
  //      a thread has all of its pieces in flight instead of paying one global-memory latency per piece
  build_q_rowtab(0);
  for (int idx = tid; idx < p.NK * WCH16; idx += NT) {
    const int r = idx / WCH16, ch = idx - r * WCH16;
    const int off = rowtab[1024 + r];
    const uint32_t so = (ch >> 3) * p.NK * 128 + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
    if (off >= 0) {
      cp_async16(aK + so, base + off + p.dim_out + ch * 8);
      cp_async16(aV + so, base + off + 2 * p.dim_out + ch * 8);
    } else {
      tc::sts128(aK + so, off == -1 ? sBias[WCH16 + ch] : make_uint4(0, 0, 0, 0));
      tc::sts128(aV + so, off == -1 ? sBias[2 * WCH16 + ch] : make_uint4(0, 0, 0, 0));
    }
  }
  __syncthreads();                                    // Q row table of tile 0
  gather_q(0);
  cp_async_commit();                                  // group: K, V, Q tile 0

  const int qtr = warp & 3, half = warp >> 2;       // TMEM lane quarter / column half handled by this warp
  const int r = qtr * 32 + lane;                    // query row inside the tile
  const uint32_t lane_addr = (uint32_t)(qtr * 32) << 16;
  // column chunks this warp's 32 rows can touch (their windows' key slots), split between the two halves
  int clo, chi;
  if (p.q_tiles > 1) { clo = 0; chi = (p.NK + 31) >> 5; }
  else {
    const int wlo = (qtr * 32) / p.lq_w;
    int whi = (qtr * 32 + 31) / p.lq_w;
    if (whi > p.G - 1) whi = p.G - 1;
    clo = (wlo * p.lk_w) >> 5;
    chi = wlo > whi ? clo : (min((whi + 1) * p.lk_w, p.NK) + 31) >> 5;
  }
  const int cmid = clo + ((chi - clo + 1) >> 1);
  const int cb = half ? cmid : clo, ce = half ? chi : cmid;          // my chunks [cb, ce), at most 4
  const int nchunks = (p.NK + 31) >> 5;

  for (int qt = 0; qt < p.q_tiles; ++qt) {
    const uint32_t ph = (uint32_t)qt & 1u;
    // ---- the next query tile's Q is fetched under this tile's MMAs and softmax
    if (qt + 1 < p.q_tiles) {
      build_q_rowtab(qt + 1);
      __syncthreads();
      gather_q(qt + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    tc::fence_proxy_async();
    __syncthreads();
    const uint32_t aQt = aQ + (uint32_t)qt * (WDCH * 128 * 128);

    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      const uint32_t idesc = tc::make_idesc_bf16(128, p.NK, 0, 0);
#pragma unroll
      for (int kk = 0; kk < WD / 16; ++kk)
        tc::umma_bf16(tS, tc::desc_kmajor_sw128(aQt + (kk >> 2) * 128 * 128 + (kk & 3) * 32),
                      tc::desc_kmajor_sw128(aK + (kk >> 2) * p.NK * 128 + (kk & 3) * 32), idesc, kk ? 1u : 0u);
      tc::umma_commit(bar_s);
    }

    // ---- softmax: all 8 warps; a thread owns row r and the column chunks [cb, ce) of it
    int wl, i;
    if (p.q_tiles > 1) { wl = 0; i = qt * 128 + r; }
    else { wl = r / p.lq_w; i = r - wl * p.lq_w; }
    const int w = w0 + wl;
    const bool valid = wl < p.G && w < nwin && i < p.lq_w;
    const int k0 = wl * p.lk_w, k1 = valid ? k0 + p.lk_w : k0;     // this row's key slots
    int out_y = -1, out_x = -1;
    if (valid) {
      const int wy = w / p.nwx, wx = w - wy * p.nwx;
      const int wo = p.qpool ? (p.ws >> 1) : p.ws;
      const int y = wy * wo + i / wo, x = wx * wo + i % wo;
      if (y < p.Ho && x < p.Wo) { out_y = y; out_x = x; }
    }
    tc::mbar_wait(bar_s, ph);
    tc::tc_fence_after();
    float mx = -INFINITY;
    for (int c = cb; c < ce; ++c) {
      uint32_t sv[32];
      tc::tmem_ld32(tS + lane_addr + c * 32, sv);
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = c * 32 + e;
        if (col >= k0 && col < k1) mx = fmaxf(mx, __uint_as_float(sv[e]));
      }
    }
    mx *= p.c;
    mxs[half * 128 + r] = mx;
    __syncthreads();
    mx = fmaxf(mx, mxs[(half ^ 1) * 128 + r]);
    // p = 2^(s*c - mx) on the row's own window, 0 elsewhere, kept in registers until BOTH halves have read S
    uint32_t pk[MAXC][16];
    float l = 0.f;
#pragma unroll
    for (int cc = 0; cc < MAXC; ++cc) {
      const int c = cb + cc;
      if (c < ce) {
        uint32_t sv[32];
        tc::tmem_ld32(tS + lane_addr + c * 32, sv);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const int col = c * 32 + e;
          float p0 = 0.f, p1 = 0.f;
          if (col >= k0 && col < k1) p0 = ex2w(fmaf(__uint_as_float(sv[e]), p.c, -mx));
          if (col + 1 >= k0 && col + 1 < k1) p1 = ex2w(fmaf(__uint_as_float(sv[e + 1]), p.c, -mx));
          l += p0 + p1;
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[cc][e >> 1] = *(uint32_t*)&hh;
        }
      }
    }
    __syncthreads();                                 // every S column has been read: P may overwrite S in place
    mxs[half * 128 + r] = l;
    // P (bf16 pairs) over the first NK/2 columns of S; chunks outside this warp pair's range are zero for these rows
    {
      uint32_t zero[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) zero[e] = 0;
      for (int c = half; c < nchunks; c += 2)
        if (c < clo || c >= chi) tc::tmem_st16(tS + lane_addr + c * 16, zero);
#pragma unroll
      for (int cc = 0; cc < MAXC; ++cc)
        if (cb + cc < ce) tc::tmem_st16(tS + lane_addr + (cb + cc) * 16, pk[cc]);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    l += mxs[(half ^ 1) * 128 + r];
    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      const uint32_t idesc = tc::make_idesc_bf16(128, WD, 0, 1);
      for (int kk = 0; kk < p.NK / 16; ++kk)
        tc::umma_bf16_ts(tO, tS + kk * 8, tc::desc_mnmajor_sw128(aV + kk * 2048, (uint32_t)p.NK * 128), idesc, kk ? 1u : 0u);
      tc::umma_commit(bar_o);
    }
    // ---- epilogue: O / l -> bf16 -> scatter; the two halves take alternate 32-column chunks
    tc::mbar_wait(bar_o, ph);
    tc::tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.out + (((long)b * p.Ho + (out_y < 0 ? 0 : out_y)) * p.Wo + (out_x < 0 ? 0 : out_x)) * p.dim_out + h * WD;
#pragma unroll 1
    for (int c = half; c < WD / 32; c += 2) {
      uint32_t o[32];
      tc::tmem_ld32(tO + lane_addr + c * 32, o);
      tc::tmem_ld_wait();
      if (out_y >= 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          uint32_t* vv = (uint32_t*)&v;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * e]) * inv,
                                                      __uint_as_float(o[g * 8 + 2 * e + 1]) * inv);
            vv[e] = *(uint32_t*)&hh;
          }
          *(uint4*)(orow + c * 32 + g * 8) = v;
        }
      }
    }
    tc::tc_fence_before();
    __syncthreads();                                 // S / O are reused by the next query tile
    tc::tc_fence_after();
  }
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

bool ms2_window_attention_tc_supported(int dt, int heads, int D, int ws, int qpool) {
  const int lk = ws * ws;
  return dt == MS2_BF16 && D == WD && lk <= 256 && ((heads * D) % 8 == 0);
}

int ms2_window_attention_tc_launch(const void* qkv, const float* qkv_bias, void* out, int B, int H, int W, int heads,
                                   int ws, int qpool, float scale, cudaStream_t st) {
  MS2_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), "window_attention_tc: 16-byte alignment");
  WinP p;
  p.qkv = (const bf16*)qkv; p.bias = qkv_bias; p.out = (bf16*)out;
  p.B = B; p.H = H; p.W = W; p.heads = heads; p.ws = ws; p.qpool = qpool;
  p.nwy = (H + ws - 1) / ws; p.nwx = (W + ws - 1) / ws;
  p.Ho = qpool ? H / 2 : H; p.Wo = qpool ? W / 2 : W;
  p.dim_out = heads * WD;
  p.lk_w = ws * ws;
  p.lq_w = qpool ? (ws / 2) * (ws / 2) : ws * ws;
  p.c = scale * 1.4426950408889634f;
  if (p.lq_w > 128) {
    p.G = 1;
    p.q_tiles = (p.lq_w + 127) / 128;
  } else {
    int g = 128 / p.lq_w, gk = 128 / p.lk_w;
    if (gk < g) g = gk;
    p.G = g < 1 ? 1 : g;
    p.q_tiles = 1;
  }
  p.NK = ((p.G * p.lk_w + 15) / 16) * 16;
  const int nwin = p.nwx * p.nwy;
  p.groups = (nwin + p.G - 1) / p.G;
  p.tmem_cols = p.NK <= 128 ? 256 : 512;
  p.o_col = p.NK <= 128 ? 128 : 256;
  MS2_CHECK_ARG(((long)H * W * 3 * p.dim_out) < (1L << 31), "window_attention_tc: token offsets must fit 31 bits");
  MS2_CHECK_ARG((p.NK + 31) / 32 <= 8, "window_attention_tc: at most 256 key slots per tile");
  const size_t smem = (size_t)p.q_tiles * WDCH * 128 * 128 + 2 * (size_t)WDCH * p.NK * 128 + 3 * WCH16 * 16 + (2 * 128 * 4 + 256) * 4 +
                      256 * 4 + 64 + 1024;
  dim3 grid(p.groups, heads, B);
  if (p.NK <= 128) {
    static size_t attr2 = 0;
    if (smem > attr2) {
      MS2_CUDA(cudaFuncSetAttribute(win_attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "win_attn_tc attr");
      attr2 = smem;
    }
    ms2_launch(win_attn_tc_kernel<2>, grid, NT, smem, st, p);
  } else {
    static size_t attr4 = 0;
    if (smem > attr4) {
      MS2_CUDA(cudaFuncSetAttribute(win_attn_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "win_attn_tc attr");
      attr4 = smem;
    }
    ms2_launch(win_attn_tc_kernel<4>, grid, NT, smem, st, p);
  }
  MS2_CHECK_LAUNCH("win_attn_tc_kernel");
  return MS2_OK;
}
