// Micro-benchmark: what blocks the tcgen05.mma issuing thread?  16-MMA batches (M=128, N=128, K=16, A in tensor memory)
// separated by (0) nothing, (1) tcgen05.fence::after_thread_sync, (2) tcgen05.commit to a local mbarrier,
// (3) four mbarrier.test_wait polls, (4) commit + fence, (5) a clock64 read + global store (the tracer's footprint).
// Reports clocks per MMA over 400 batches and the time the issuing thread spends per batch (issue-side clock).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I medical-sam2_b200/medsam2_b200/csrc -o tools/ubench/umma_issue tools/ubench/umma_issue.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int rounds, long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2[2];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&bar2[0], 1); tc::mbar_init(&bar2[1], 1); tc::fence_barrier_init(); }
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i % 7;
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, 128, 0, 0);
    const uint32_t aB = tc::smem_u32(smem);
    if (tc::elect_one()) {
      long long t0 = clock64();
      uint32_t polls = 0;
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint64_t db = tc::desc_kmajor_sw128(aB + (r & 1) * 65536 + (kk >> 2) * 128 * 128 + (kk & 3) * 32);
          tc::umma_bf16_ts(tb + 128 + (r & 1) * 128, tb + 384 + kk * 8, db, idesc, 1u);
        }
        if (MODE == 1 || MODE == 4) tc::tc_fence_after();
        if (MODE == 2 || MODE == 4) tc::umma_commit(&bar2[r & 1]);
        if (MODE == 3) for (int q = 0; q < 4; ++q) polls += tc::mbar_test_wait(&bar2[q & 1], 0);
        if (MODE == 5) sink[1 + (r & 7)] = clock64();
      }
      long long t_issue = clock64();
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      long long t1 = clock64();
      cyc[blockIdx.x] = t1 - t0;
      cyc[148 + blockIdx.x] = t_issue - t0;
      if (polls == 0xffffffffu) sink[0] = polls;
    }
    __syncwarp();
  } else if (MODE >= 6 && warp <= 2) {
    // background: one or two OTHER threads spin on mbarrier.test_wait (6: one spinner, 7: two, 8: two with try_wait)
    if (tc::elect_one() && (warp == 1 || MODE >= 7)) {
      uint32_t polls = 0;
      while (!*(volatile int*)&sink[16]) {        // never set: bounded by the iteration count below
        polls += MODE == 8 ? tc::mbar_try_wait(&bar2[warp & 1], 0) : tc::mbar_test_wait(&bar2[warp & 1], 0);
        if (tc::mbar_test_wait(&bar, 0)) break;   // the MMA warp's final commit has landed
      }
      if (polls == 0xffffffffu) sink[0] = polls;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int MODE>
void run(const char* name) {
  long long *cyc, *sink; cudaMalloc(&cyc, 2 * 148 * 8); cudaMalloc(&sink, 256); cudaMemset(sink, 0, 256);
  const int rounds = 400, smem = 200 * 1024, ctas = 148;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE><<<ctas, 128, smem>>>(cyc, rounds, sink);
  k<MODE><<<ctas, 128, smem>>>(cyc, rounds, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0, ci = 0; for (int i = 0; i < ctas; ++i) { c += h[i]; ci += h[148 + i]; } c /= ctas; ci /= ctas;
  printf("%-46s %.1f clk per MMA end to end, issue loop %.1f clk per MMA  %s\n", name, c / (rounds * 16.0), ci / (rounds * 16.0),
         cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<0>("16-MMA batches back to back");
  run<1>("+ tcgen05.fence::after_thread_sync per batch");
  run<2>("+ tcgen05.commit per batch");
  run<3>("+ 4 mbarrier.test_wait per batch");
  run<4>("+ commit + fence per batch");
  run<5>("+ clock64 + st.global per batch");
  run<6>("one other thread spinning on test_wait");
  run<7>("two other threads spinning on test_wait");
  run<8>("two other threads spinning on try_wait");
  return 0;
}
