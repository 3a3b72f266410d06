// tcgen05/TMEM/TMA bf16 GEMM — placeholder until phase 2 lands (declares "unsupported" so that
// ms2_gemm(impl=0) routes everything to the SIMT kernel).
#include "common.cuh"

bool ms2_gemm_tc_supported(int a_dt, int w_dt, long lda, long ldo, int M, int N, int K) { return false; }

int ms2_gemm_tc_launch(const void* A, long lda, const void* W, const float* bias, const float* colscale,
                       const float* residual, long ldr, void* out, int o_dt, long ldo, int M, int N, int K,
                       int act, cudaStream_t st) {
  ms2_set_error("gemm_tc: not built");
  return MS2_ERR_UNSUPPORTED;
}
