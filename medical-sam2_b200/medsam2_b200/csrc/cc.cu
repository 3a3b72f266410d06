// 8-connectivity connected components + fused hole filling for sm_100a.
//
// Replaces the reference's only native code (sam2_train/csrc/connected_components.cu:62-282):
// six kernels x N images in a host loop + three zero-filled N*H*W int32 tensors.  Here one CTA owns
// one image: the 2x2-block union-find lives in shared memory (a 256x256 mask = 16384 blocks = 64 KiB
// of labels + 64 KiB of areas), all N images run in ONE launch, nothing is zero-filled in HBM and
// the hole-filling `where` (utils/misc.py:247-258) can be fused so labels never touch HBM at all.
// Label semantics are the reference's: 1 + smallest (row&~1)*W+(col&~1) over the component.
// Large images (more blocks than fit in shared memory) take a global-memory path of four launches
// for the whole batch.
#include "common.cuh"

namespace {

constexpr int kMaxSmemBlocks = 28672;  // 2 * 4 B * 28672 = 224 KiB of dynamic shared memory

__device__ __forceinline__ int uf_find(volatile int* lab, int n) {
  while (lab[n] != n) n = lab[n];
  return n;
}

__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
  bool done;
  do {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a < b) {
      int old = atomicMin(lab + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(lab + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// foreground test for the two input flavours
template <bool SCORES> struct Src;
template <> struct Src<false> {
  const uint8_t* p;
  float thresh;
  __device__ __forceinline__ bool fg(long i) const { return p[i] != 0; }
};
template <> struct Src<true> {
  const float* p;
  float thresh;
  __device__ __forceinline__ bool fg(long i) const { return p[i] <= thresh; }
};

// bits of a 2x2 block: 1=(r,c) 2=(r,c+1) 4=(r+1,c) 8=(r+1,c+1)
template <bool SCORES>
__device__ __forceinline__ unsigned block_bits(const Src<SCORES>& s, long base, int W) {
  unsigned b = 0;
  if (s.fg(base)) b |= 1u;
  if (s.fg(base + 1)) b |= 2u;
  if (s.fg(base + W)) b |= 4u;
  if (s.fg(base + W + 1)) b |= 8u;
  return b;
}

// One CTA per image; labels/areas of the 2x2 blocks in shared memory.
template <bool SCORES>
__global__ void __launch_bounds__(1024, 1)
cc_smem_kernel(Src<SCORES> src, int32_t* __restrict__ labels, int32_t* __restrict__ counts,
               float* __restrict__ filled, int H, int W, int max_area, float fill_value) {
  MS2_PDL_WAIT();
  extern __shared__ int smem[];
  const int BW = W >> 1, BH = H >> 1, NB = BW * BH;
  int* lab = smem;
  int* area = smem + NB;
  unsigned char* bits = reinterpret_cast<unsigned char*>(smem + 2 * NB);
  const long img = (long)blockIdx.x * H * W;
  Src<SCORES> s = src;
  s.p += img;

  for (int b = threadIdx.x; b < NB; b += blockDim.x) {
    int by = b / BW, bx = b - by * BW;
    bits[b] = (unsigned char)block_bits(s, (long)(2 * by) * W + 2 * bx, W);
    lab[b] = b;
    area[b] = 0;
  }
  __syncthreads();
  // merge with the TL / T / TR / L neighbour blocks (8-connectivity between their border pixels)
  for (int b = threadIdx.x; b < NB; b += blockDim.x) {
    unsigned me = bits[b];
    if (!me) continue;
    int by = b / BW, bx = b - by * BW;
    if (by > 0) {
      unsigned t = bits[b - BW];
      if ((me & 3u) && (t & 12u)) uf_union(lab, b, b - BW);
      if (bx > 0 && (me & 1u) && (bits[b - BW - 1] & 8u)) uf_union(lab, b, b - BW - 1);
      if (bx + 1 < BW && (me & 2u) && (bits[b - BW + 1] & 4u)) uf_union(lab, b, b - BW + 1);
    }
    if (bx > 0 && (me & 5u) && (bits[b - 1] & 10u)) uf_union(lab, b, b - 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < NB; b += blockDim.x) {
    if (!bits[b]) continue;
    int r = uf_find(lab, b);
    lab[b] = r;  // roots never change after the merge phase, so racing compressions are benign
    atomicAdd(area + r, __popc((unsigned)bits[b]));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < NB; b += blockDim.x) {
    int by = b / BW, bx = b - by * BW;
    unsigned me = bits[b];
    int r = me ? lab[b] : 0;
    int a = me ? area[r] : 0;
    int rby = r / BW, rbx = r - rby * BW;
    int lv = (2 * rby) * W + 2 * rbx + 1;
    long o = img + (long)(2 * by) * W + 2 * bx;
    if (filled) {
      bool hole = me && a <= max_area;
      const float* in = reinterpret_cast<const float*>(src.p) + o;
      float2 r0 = *reinterpret_cast<const float2*>(in);
      float2 r1 = *reinterpret_cast<const float2*>(in + W);
      if (hole) {
        if (me & 1u) r0.x = fill_value;
        if (me & 2u) r0.y = fill_value;
        if (me & 4u) r1.x = fill_value;
        if (me & 8u) r1.y = fill_value;
      }
      *reinterpret_cast<float2*>(filled + o) = r0;
      *reinterpret_cast<float2*>(filled + o + W) = r1;
    } else {
      int2 l0 = make_int2((me & 1u) ? lv : 0, (me & 2u) ? lv : 0);
      int2 l1 = make_int2((me & 4u) ? lv : 0, (me & 8u) ? lv : 0);
      int2 c0 = make_int2((me & 1u) ? a : 0, (me & 2u) ? a : 0);
      int2 c1 = make_int2((me & 4u) ? a : 0, (me & 8u) ? a : 0);
      *reinterpret_cast<int2*>(labels + o) = l0;
      *reinterpret_cast<int2*>(labels + o + W) = l1;
      *reinterpret_cast<int2*>(counts + o) = c0;
      *reinterpret_cast<int2*>(counts + o + W) = c1;
    }
  }
}

// ---------------- global-memory path (large images); the union-find array is `ws` indexed by block id
__global__ void ccg_init(const uint8_t* __restrict__ mask, int32_t* __restrict__ ws, int H, int W) {
  MS2_PDL_WAIT();
  const int BW = W >> 1, NB = BW * (H >> 1);
  const long img = (long)blockIdx.z * H * W;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  ws[img + b] = b;
  ws[img + NB + b] = 0;  // areas
}
__device__ __forceinline__ unsigned g_bits(const uint8_t* m, int by, int bx, int W) {
  Src<false> s{m, 0.f};
  return block_bits(s, (long)(2 * by) * W + 2 * bx, W);
}
__global__ void ccg_merge(const uint8_t* __restrict__ mask, int32_t* __restrict__ ws, int H, int W) {
  MS2_PDL_WAIT();
  const int BW = W >> 1, NB = BW * (H >> 1);
  const long img = (long)blockIdx.z * H * W;
  const uint8_t* m = mask + img;
  int* lab = ws + img;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  int by = b / BW, bx = b - by * BW;
  unsigned me = g_bits(m, by, bx, W);
  if (!me) return;
  if (by > 0) {
    if ((me & 3u) && (g_bits(m, by - 1, bx, W) & 12u)) uf_union(lab, b, b - BW);
    if (bx > 0 && (me & 1u) && (g_bits(m, by - 1, bx - 1, W) & 8u)) uf_union(lab, b, b - BW - 1);
    if (bx + 1 < BW && (me & 2u) && (g_bits(m, by - 1, bx + 1, W) & 4u)) uf_union(lab, b, b - BW + 1);
  }
  if (bx > 0 && (me & 5u) && (g_bits(m, by, bx - 1, W) & 10u)) uf_union(lab, b, b - 1);
}
__global__ void ccg_compress_count(const uint8_t* __restrict__ mask, int32_t* __restrict__ ws, int H, int W) {
  MS2_PDL_WAIT();
  const int BW = W >> 1, NB = BW * (H >> 1);
  const long img = (long)blockIdx.z * H * W;
  int* lab = ws + img;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  int by = b / BW, bx = b - by * BW;
  unsigned me = g_bits(mask + img, by, bx, W);
  if (!me) return;
  int r = uf_find(lab, b);
  lab[b] = r;
  atomicAdd(lab + NB + r, __popc(me));
}
__global__ void ccg_final(const uint8_t* __restrict__ mask, const int32_t* __restrict__ ws,
                          int32_t* __restrict__ labels, int32_t* __restrict__ counts, int H, int W) {
  MS2_PDL_WAIT();
  const int BW = W >> 1, NB = BW * (H >> 1);
  const long img = (long)blockIdx.z * H * W;
  const int* lab = ws + img;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  int by = b / BW, bx = b - by * BW;
  unsigned me = g_bits(mask + img, by, bx, W);
  int r = me ? lab[b] : 0;
  int a = me ? lab[NB + r] : 0;
  int rby = r / BW, rbx = r - rby * BW;
  int lv = (2 * rby) * W + 2 * rbx + 1;
  long o = img + (long)(2 * by) * W + 2 * bx;
  *reinterpret_cast<int2*>(labels + o) = make_int2((me & 1u) ? lv : 0, (me & 2u) ? lv : 0);
  *reinterpret_cast<int2*>(labels + o + W) = make_int2((me & 4u) ? lv : 0, (me & 8u) ? lv : 0);
  *reinterpret_cast<int2*>(counts + o) = make_int2((me & 1u) ? a : 0, (me & 2u) ? a : 0);
  *reinterpret_cast<int2*>(counts + o + W) = make_int2((me & 4u) ? a : 0, (me & 8u) ? a : 0);
}

size_t smem_bytes(int NB) { return (size_t)NB * 8 + (size_t)((NB + 3) & ~3); }

}  // namespace

extern "C" int ms2_cc_label(const uint8_t* mask, int32_t* labels, int32_t* counts, int32_t* workspace,
                            int N, int H, int W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MS2_CHECK_ARG(N >= 0 && H > 0 && W > 0, "cc_label: bad shape");
  MS2_CHECK_ARG((H % 2) == 0, "height must be a even number");
  MS2_CHECK_ARG((W % 2) == 0, "width must be a even number");
  if (N == 0) return MS2_OK;
  MS2_CHECK_ARG(mask && labels && counts, "cc_label: null pointer");
  const int NB = (H / 2) * (W / 2);
  if (NB <= kMaxSmemBlocks) {
    size_t sm = smem_bytes(NB);
    MS2_CUDA(cudaFuncSetAttribute(cc_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm),
             "cc_label attr");
    Src<false> s{mask, 0.f};
    ms2_launch(cc_smem_kernel<false>, N, 1024, sm, stream, s, labels, counts, nullptr, H, W, 0, 0.f);
    MS2_CHECK_LAUNCH("cc_smem_kernel");
    return MS2_OK;
  }
  MS2_CHECK_ARG(workspace != nullptr, "cc_label: workspace required for %dx%d", H, W);
  dim3 grid(ceil_div(NB, 256), 1, N);
  ms2_launch(ccg_init, grid, 256, 0, stream, mask, workspace, H, W);
  ms2_launch(ccg_merge, grid, 256, 0, stream, mask, workspace, H, W);
  ms2_launch(ccg_compress_count, grid, 256, 0, stream, mask, workspace, H, W);
  ms2_launch(ccg_final, grid, 256, 0, stream, mask, workspace, labels, counts, H, W);
  MS2_CHECK_LAUNCH("cc global path");
  return MS2_OK;
}

namespace {
// Hole filling only asks "is this background pixel in an 8-connected component of area <= max_area?".  For the
// small areas used on the hot path (fill_hole_area = 8) that is a LOCAL question: a bounded flood fill from the
// pixel either exhausts its component within max_area pixels (hole) or finds a (max_area+1)-th pixel (not a
// hole).  One thread per pixel, at most (max_area+1)*8 neighbour probes out of L1/L2 - no union-find, no global
// labels, and bit-identical to labelling the whole image (utils/misc.py:247-258).
constexpr int kLocalMaxArea = 32;
__global__ void __launch_bounds__(256)
fill_holes_local_kernel(const float* __restrict__ in, float* __restrict__ out, long total, int H, int W, float thresh,
                        int max_area, float fill_value) {
  MS2_PDL_WAIT();
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int hw = H * W;
  const long n = p / hw;
  const int pix = (int)(p - n * hw), y0 = pix / W, x0 = pix - y0 * W;
  const float* img = in + n * hw;
  const float v = img[pix];
  float res = v;
  if (v <= thresh) {
    // every background pixel of the 3x3 neighbourhood is an 8-neighbour of this one, i.e. in its component: more than
    // max_area of them settle the question without a flood fill (almost every pixel of a real mask: max_area = 8 < 9)
    int near = 1;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int dy = (d < 3) ? -1 : (d < 5 ? 0 : 1);
      const int dx = (d == 0 || d == 3 || d == 5) ? -1 : ((d == 1 || d == 6) ? 0 : 1);
      const int ny = y0 + dy, nx = x0 + dx;
      if (ny >= 0 && ny < H && nx >= 0 && nx < W) near += (img[ny * W + nx] <= thresh) ? 1 : 0;
    }
    if (near > max_area) {
      out[p] = v;
      return;
    }
    int q[kLocalMaxArea];
    q[0] = (y0 << 16) | x0;
    int cnt = 1, head = 0;
    bool small = true;
    while (small && head < cnt) {
      const int cy = q[head] >> 16, cx = q[head] & 0xffff;
      ++head;
#pragma unroll 1
      for (int d = 0; d < 8 && small; ++d) {
        const int dy = (d < 3) ? -1 : (d < 5 ? 0 : 1);
        const int dx = (d == 0 || d == 3 || d == 5) ? -1 : ((d == 1 || d == 6) ? 0 : 1);
        const int ny = cy + dy, nx = cx + dx;
        if (ny < 0 || ny >= H || nx < 0 || nx >= W) continue;
        if (!(img[ny * W + nx] <= thresh)) continue;
        const int key = (ny << 16) | nx;
        bool seen = false;
        for (int i = 0; i < cnt; ++i) seen |= (q[i] == key);
        if (seen) continue;
        if (cnt == max_area) { small = false; break; }
        q[cnt++] = key;
      }
    }
    if (small) res = fill_value;
  }
  out[p] = res;
}
}  // namespace

extern "C" int ms2_fill_holes(const float* in, float* out, int N, int H, int W, float thresh, int max_area,
                              float fill_value, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MS2_CHECK_ARG((H % 2) == 0 && (W % 2) == 0 && H > 0 && W > 0, "fill_holes: H, W must be even");
  MS2_CHECK_ARG(max_area > 0, "max_area must be positive");
  if (N == 0) return MS2_OK;
  MS2_CHECK_ARG(in && out, "fill_holes: null pointer");
  if (max_area <= kLocalMaxArea && H < 65536 && W < 65536) {
    const long total = (long)N * H * W;
    ms2_launch(fill_holes_local_kernel, ceil_div(total, 256), 256, 0, stream, in, out, total, H, W, thresh, max_area, fill_value);
    MS2_CHECK_LAUNCH("fill_holes_local_kernel");
    return MS2_OK;
  }
  const int NB = (H / 2) * (W / 2);
  MS2_CHECK_ARG(NB <= kMaxSmemBlocks, "fill_holes: %dx%d exceeds the shared-memory path", H, W);
  size_t sm = smem_bytes(NB);
  MS2_CUDA(cudaFuncSetAttribute(cc_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm),
           "fill_holes attr");
  Src<true> s{in, thresh};
  ms2_launch(cc_smem_kernel<true>, N, 1024, sm, stream, s, nullptr, nullptr, out, H, W, max_area, fill_value);
  MS2_CHECK_LAUNCH("cc_smem_kernel<fill>");
  return MS2_OK;
}
