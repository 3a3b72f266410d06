// Hiera windowed attention on tcgen05 (bf16, head dim 96): hieradet.py:58-83,136-159 +
// backbones/utils.py:16-62.  Window partition, zero-pad-AFTER-norm semantics (pad tokens enter softmax
// with k = b_k, v = b_v: SURVEY §0 finding 6), the 2x2 q max-pool and unpartition+crop are folded into the
// gather that builds the swizzled shared-memory operands and into the scatter of the epilogue, so no padded
// or re-ordered copy of the tokens ever exists in HBM.
//
// One CTA = one group of G windows of one head packed into a single M=128 tile (block-diagonal mask):
//   all warps : gather Q (<=128 rows, optional 2x2 max-pool), K, V (NK <= 256 key slots) -> SW128 smem
//   thread 0  : S = Q K^T  (tcgen05.mma M=128, N=NK, K=96)                       -> TMEM
//   warps 0-3 : row softmax restricted to the row's own window -> P (bf16, swizzled smem, aliases Q/K)
//   thread 0  : O = P V    (M=128, N=96, K=NK; V is MN-major)                    -> TMEM
//   warps 0-3 : O / l -> bf16 -> scatter to [B,Ho,Wo,heads*D]
// Windows of 196 tokens (stage 3) use two query tiles per window.  Small-window configurations need
// <= 113 KB of shared memory and 256 TMEM columns, so two CTAs share an SM and overlap each other's phases.
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

__global__ void __launch_bounds__(NT)
win_attn_tc_kernel(const WinP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kv_bytes = WDCH * p.NK * 128;
  uint8_t* sQ = smem;                               // 2 x [128][128 B]
  uint8_t* sK = sQ + WDCH * 128 * 128;              // 2 x [NK][128 B]
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sP = smem;                               // aliases Q (+K): ceil(NK/64) x [128][128 B]
  uint8_t* tail = sV + kv_bytes;
  uint4* sBias = (uint4*)tail;                      // [3][12] 16-byte chunks of the bf16 bias of this head
  uint64_t* bar_s = (uint64_t*)(tail + 3 * WCH16 * 16);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_ptr = (uint32_t*)(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.x / p.q_tiles, qt = blockIdx.x - grp * p.q_tiles;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nwin = p.nwx * p.nwy;
  const int w0 = grp * p.G;
  const long C3 = 3L * p.dim_out;
  const bf16* base = p.qkv + (long)b * p.H * p.W * C3;

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
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // value of q/k/v (which) chunk ch at padded-grid position (y,x): pad -> bias
  auto tok = [&](int which, int y, int x, int ch) -> uint4 {
    if (y < p.H && x < p.W)
      return __ldg((const uint4*)(base + ((long)y * p.W + x) * C3 + which * p.dim_out + h * WD + ch * 8));
    return sBias[which * WCH16 + ch];
  };
  auto swz = [](uint8_t* chunk0, int rows, int row, int ch) -> uint4* {
    return (uint4*)(chunk0 + (ch >> 3) * rows * 128 + row * 128 + (((ch & 7) ^ (row & 7)) << 4));
  };

  // ---- gather Q
  for (int idx = tid; idx < 128 * WCH16; idx += NT) {
    const int r = idx / WCH16, ch = idx - r * WCH16;
    uint4 v = make_uint4(0, 0, 0, 0);
    int wl, i;
    if (p.q_tiles > 1) { wl = 0; i = qt * 128 + r; }
    else { wl = r / p.lq_w; i = r - wl * p.lq_w; }
    const int w = w0 + wl;
    if (wl < p.G && w < nwin && i < p.lq_w) {
      const int wy = w / p.nwx, wx = w - wy * p.nwx;
      if (!p.qpool) {
        v = tok(0, wy * p.ws + i / p.ws, wx * p.ws + i % p.ws, ch);
      } else {
        const int hw = p.ws >> 1;
        const int y = wy * p.ws + 2 * (i / hw), x = wx * p.ws + 2 * (i % hw);
        v = max_bf16x8(max_bf16x8(tok(0, y, x, ch), tok(0, y, x + 1, ch)),
                       max_bf16x8(tok(0, y + 1, x, ch), tok(0, y + 1, x + 1, ch)));
      }
    }
    *swz(sQ, 128, r, ch) = v;
  }
  // ---- gather K, V
  for (int idx = tid; idx < p.NK * WCH16; idx += NT) {
    const int r = idx / WCH16, ch = idx - r * WCH16;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    const int wl = r / p.lk_w, j = r - wl * p.lk_w;
    const int w = w0 + wl;
    if (wl < p.G && w < nwin) {
      const int wy = w / p.nwx, wx = w - wy * p.nwx;
      const int y = wy * p.ws + j / p.ws, x = wx * p.ws + j % p.ws;
      kv = tok(1, y, x, ch);
      vv = tok(2, y, x, ch);
    }
    *swz(sK, p.NK, r, ch) = kv;
    *swz(sV, p.NK, r, ch) = vv;
  }
  tc::fence_proxy_async();
  __syncthreads();

  const uint32_t tS = tmem_base, tO = tmem_base + (uint32_t)p.o_col;
  if (warp == 0 && tc::elect_one()) {
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_bf16(128, p.NK, 0, 0);
    const uint32_t aQ = tc::smem_u32(sQ), aK = tc::smem_u32(sK);
#pragma unroll
    for (int kk = 0; kk < WD / 16; ++kk)
      tc::umma_bf16(tS, tc::desc_kmajor_sw128(aQ + (kk >> 2) * 128 * 128 + (kk & 3) * 32),
                    tc::desc_kmajor_sw128(aK + (kk >> 2) * p.NK * 128 + (kk & 3) * 32), idesc, kk ? 1u : 0u);
    tc::umma_commit(bar_s);
  }

  float l = 0.f;
  int out_y = -1, out_x = -1;
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    int wl, i;
    if (p.q_tiles > 1) { wl = 0; i = qt * 128 + r; }
    else { wl = r / p.lq_w; i = r - wl * p.lq_w; }
    const int w = w0 + wl;
    const bool valid = wl < p.G && w < nwin && i < p.lq_w;
    const int k0 = wl * p.lk_w, k1 = valid ? k0 + p.lk_w : k0;     // this row's key slots
    if (valid) {
      const int wy = w / p.nwx, wx = w - wy * p.nwx;
      const int wo = p.qpool ? (p.ws >> 1) : p.ws;
      const int y = wy * wo + i / wo, x = wx * wo + i % wo;
      if (y < p.Ho && x < p.Wo) { out_y = y; out_x = x; }
    }
    tc::mbar_wait(bar_s, 0);
    tc::tc_fence_after();
    // pass 1: row max over own window
    float mx = -INFINITY;
    for (int c0 = 0; c0 < p.NK; c0 += 32) {
      uint32_t s[32];
      tc::tmem_ld32(tS + lane_addr + c0, s);
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = c0 + e;
        if (col >= k0 && col < k1) mx = fmaxf(mx, __uint_as_float(s[e]) * p.c);
      }
    }
    // pass 2: p = 2^(s*c - mx) on own window, 0 elsewhere -> swizzled P (K-major, chunks of 64 keys)
    for (int c0 = 0; c0 < p.NK; c0 += 32) {
      uint32_t s[32];
      tc::tmem_ld32(tS + lane_addr + c0, s);
      tc::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int col = c0 + e;
        float p0 = 0.f, p1 = 0.f;
        if (col >= k0 && col < k1) p0 = ex2w(__uint_as_float(s[e]) * p.c - mx);
        if (col + 1 >= k0 && col + 1 < k1) p1 = ex2w(__uint_as_float(s[e + 1]) * p.c - mx);
        l += p0 + p1;
        __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
        pk[e >> 1] = *(uint32_t*)&hh;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = c0 + g * 8;
        if (col < p.NK) {
          const int kc = col >> 6, gg = (col & 63) >> 3;
          *(uint4*)(sP + kc * 128 * 128 + r * 128 + ((gg ^ (r & 7)) << 4)) =
              make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        }
      }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0 && tc::elect_one()) {
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_bf16(128, WD, 0, 1);
    const uint32_t aP = tc::smem_u32(sP), aV = tc::smem_u32(sV);
    for (int kk = 0; kk < p.NK / 16; ++kk)
      tc::umma_bf16(tO, tc::desc_kmajor_sw128(aP + (kk >> 2) * 128 * 128 + (kk & 3) * 32),
                    tc::desc_mnmajor_sw128(aV + kk * 2048, (uint32_t)p.NK * 128), idesc, kk ? 1u : 0u);
    tc::umma_commit(bar_o);
  }
  if (warp < 4) {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    tc::mbar_wait(bar_o, 0);
    tc::tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.out + (((long)b * p.Ho + (out_y < 0 ? 0 : out_y)) * p.Wo + (out_x < 0 ? 0 : out_x)) * p.dim_out + h * WD;
#pragma unroll 1
    for (int c = 0; c < WD / 32; ++c) {
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
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
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
  const int pch = (p.NK + 63) / 64;
  MS2_CHECK_ARG(pch * 128 * 128 <= WDCH * 128 * 128 + WDCH * p.NK * 128, "window_attention_tc: P does not fit");
  const size_t smem = (size_t)WDCH * 128 * 128 + 2 * (size_t)WDCH * p.NK * 128 + 3 * WCH16 * 16 + 64 + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MS2_CUDA(cudaFuncSetAttribute(win_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "win_attn_tc attr");
    attr_smem = smem;
  }
  dim3 grid(p.groups * p.q_tiles, heads, B);
  win_attn_tc_kernel<<<grid, NT, smem, st>>>(p);
  MS2_CHECK_LAUNCH("win_attn_tc_kernel");
  return MS2_OK;
}
