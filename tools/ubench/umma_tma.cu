// Micro-benchmark: tcgen05.mma rate (M=128, N=128, K=16, A in tensor memory, B K-major SW128 in shared memory) while a
// producer warp streams 64 KB tiles into OTHER shared-memory slots with TMA (cp.async.bulk.tensor), as the attention
// kernel's K/V producer does.  MODE 0: no loads; 1: plain TMA loads; 2: loads throttled to one tile per ~2000 clk.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I medical-sam2_b200/medsam2_b200/csrc -o tools/ubench/umma_tma tools/ubench/umma_tma.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, long long* cyc, int rounds, int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, full[2];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::fence_barrier_init(); done = 0; }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i % 7;
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
    const uint32_t aB = tc::smem_u32(smem);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint64_t db = tc::desc_kmajor_sw128(aB + (kk >> 2) * 128 * 128 + (kk & 3) * 32);
          tc::umma_bf16_ts(tb + 128 + (r & 1) * 128, tb + 384 + kk * 8, db, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; done = 1; }
  } else if (warp == 1 && MODE > 0) {
    if (tc::elect_one()) {
      long n = 0;
      long long next = clock64();
      while (!done) {
        const int s = n & 1;
        if (n >= 2) tc::mbar_wait(&full[s], (uint32_t)((n - 2) >> 1) & 1u);
        if (MODE == 2) { while (clock64() < next) {} next += 2000; }
        tc::mbar_arrive_expect_tx(&full[s], 65536);
        const int row = (int)(((long)blockIdx.x * 131 + n) % ntiles) * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tc::tma_load_4d(smem + 65536 + s * 65536 + c * 128 * 128, &tm, &full[s], c * 64, row, 0, 0);
        ++n;
      }
      if (n >= 1) tc::mbar_wait(&full[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1u);
      if (n >= 2) tc::mbar_wait(&full[(n - 2) & 1], (uint32_t)((n - 2) >> 1) & 1u);
      cyc[148 + blockIdx.x] = n;
    }
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int MODE>
void run(const char* name, const CUtensorMap& tm, int ntiles) {
  long long* cyc; cudaMalloc(&cyc, 2 * 148 * 8); cudaMemset(cyc, 0, 2 * 148 * 8);
  const int rounds = 400, smem = 200 * 1024, ctas = 148;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE><<<ctas, 128, smem>>>(tm, cyc, rounds, ntiles);
  k<MODE><<<ctas, 128, smem>>>(tm, cyc, rounds, ntiles);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0, it = 0; for (int i = 0; i < ctas; ++i) { c += h[i]; it += h[148 + i]; } c /= ctas; it /= ctas;
  printf("%-40s %.1f clk per MMA (full rate 64); 64 KB tiles loaded per 2048 clk: %.2f  %s\n", name, c / (rounds * 16.0),
         it / (c / 2048.0), cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  const int Lk = 209120;
  void* kbuf; cudaMalloc(&kbuf, (size_t)Lk * 256 * 2); cudaMemset(kbuf, 0x3c, (size_t)Lk * 256 * 2);
  CUtensorMap tm;
  const uint64_t dims[4] = {256, (uint64_t)Lk, 1, 1};
  const uint64_t str[3] = {256, (uint64_t)256 * Lk, (uint64_t)256 * Lk};
  const uint32_t box[4] = {64, 128, 1, 1};
  if (tc::make_tmap_bf16(&tm, kbuf, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap failed\n"); return 1; }
  run<0>("no loads", tm, Lk / 128);
  run<1>("TMA loads back to back (2-slot ring)", tm, Lk / 128);
  run<2>("TMA loads, one 64 KB tile per 2000 clk", tm, Lk / 128);
  return 0;
}
