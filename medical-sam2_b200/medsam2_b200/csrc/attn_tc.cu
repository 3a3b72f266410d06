// tcgen05 flash attention — placeholder until phase 2 lands.
#include "common.cuh"
bool ms2_attention_tc_supported(int dt, long q_hs, long q_ts, long k_ts, long v_ts, long o_ts, int Hh, int Lq,
                                int Lk, int D) { return false; }
int ms2_attention_tc_launch(const void* q, const void* k, const void* v, void* o, long q_ts, long k_ts, long v_ts,
                            long o_ts, long q_bs, long k_bs, long v_bs, long o_bs, int B, int Lq, int Lk, int D,
                            float scale, cudaStream_t st) {
  ms2_set_error("attention_tc: not built");
  return MS2_ERR_UNSUPPORTED;
}
