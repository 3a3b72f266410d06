// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) by N and by where the A operand lives.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I medical-sam2_b200/medsam2_b200/csrc -o tools/ubench/umma tools/ubench/umma.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

// A_TMEM: A from tensor memory (columns 256..), else from shared memory.  B: K-major (b_mn = 0) or MN-major tile.
template <int N, bool A_TMEM, int B_MN>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int rounds) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i % 7;
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N, 0, B_MN);
    const uint32_t aA = tc::smem_u32(smem), aB = aA + 32768;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint64_t db = B_MN ? tc::desc_mnmajor_sw128(aB + (kk & 7) * 2048, 128 * 128)
                                   : tc::desc_kmajor_sw128(aB + (kk >> 2) * N * 128 + (kk & 3) * 32);
          if (A_TMEM) tc::umma_bf16_ts(tb, tb + 256 + kk * 8, db, idesc, 1u);
          else tc::umma_bf16(tb, tc::desc_kmajor_sw128(aA + (kk >> 2) * 128 * 128 / 4 + (kk & 3) * 32), db, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int N, bool A_TMEM, int B_MN>
void run(const char* name, int ctas, int rounds = 400) {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k<N, A_TMEM, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, A_TMEM, B_MN><<<ctas, 128, smem>>>(cyc, rounds);
  k<N, A_TMEM, B_MN><<<ctas, 128, smem>>>(cyc, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < ctas; ++i) c += h[i]; c /= ctas;
  const double per = c / (rounds * 16.0);
  printf("%-34s ctas %3d rounds %3d: %.1f clk per MMA  (full rate = %d)  total %.0f clk  %s\n", name, ctas, rounds, per, N / 2, c,
         cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  for (int ctas : {1, 148}) {
    run<64, true, 0>("A tmem, N=64,  B K-major", ctas);
    run<128, true, 0>("A tmem, N=128, B K-major", ctas);
    run<256, true, 0>("A tmem, N=256, B K-major", ctas);
    run<64, true, 1>("A tmem, N=64,  B MN-major (P V)", ctas);
    run<128, true, 1>("A tmem, N=128, B MN-major", ctas);
    run<64, false, 0>("A smem, N=64,  B K-major", ctas);
    run<128, false, 0>("A smem, N=128, B K-major", ctas);
    run<256, false, 0>("A smem, N=256, B K-major", ctas);
    run<64, false, 1>("A smem, N=64,  B MN-major", ctas);
  }
  // latency of issue -> completion -> commit -> mbarrier wake-up: 16, 32, 64 MMAs then one commit
  for (int r : {1, 2, 4}) run<128, true, 0>("A tmem, N=128 (latency)", 148, r);
  for (int r : {1, 2}) run<64, true, 1>("A tmem, N=64 MN-major (latency)", 148, r);
  return 0;
}
