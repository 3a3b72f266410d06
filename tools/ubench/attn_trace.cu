// Timeline of one CTA of the memory cross-attention kernel (attn_mc_kernel): clock64 stamps of the producer, the MMA
// issuer and two softmax warps for tiles 16..47 of split 0.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMS2_ATTN_TRACE -I medical-sam2_b200/medsam2_b200/csrc \
//        -I include -o tools/ubench/attn_trace tools/ubench/attn_trace.cu medical-sam2_b200/medsam2_b200/csrc/common.cu -lcuda
#include "attn_tc.cu"
#include <vector>
#include <cstdlib>

int main() {
  const int Lq = 4096, Lk = 209120;
  std::vector<uint16_t> hq((size_t)Lq * 256), hk((size_t)Lk * 256), hv((size_t)Lk * 64);
  auto rnd = [](std::vector<uint16_t>& v) {
    for (auto& x : v) { float f = (float)rand() / RAND_MAX * 2.f - 1.f; uint32_t u; memcpy(&u, &f, 4); x = (uint16_t)(u >> 16); }
  };
  rnd(hq); rnd(hk); rnd(hv);
  void *q, *k, *v, *o, *ws;
  const long ws_bytes = 256L << 20;
  cudaMalloc(&q, hq.size() * 2); cudaMalloc(&k, hk.size() * 2); cudaMalloc(&v, hv.size() * 2);
  cudaMalloc(&o, (size_t)Lq * 64 * 2); cudaMalloc(&ws, ws_bytes);
  cudaMemcpy(q, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(k, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(v, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice);
  for (int it = 0; it < 3; ++it) {
    int rc = ms2_attention_tc_launch(q, k, v, o, (long)Lq * 256, 256, 256, (long)Lk * 256, 256, 256, (long)Lk * 64, 64, 64,
                                     (long)Lq * 64, 64, 64, 1, 1, Lq, Lk, 256, 64, 0.0625f, ws, ws_bytes, 0);
    if (rc) { printf("launch failed %d\n", rc); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[32 * 16];
  cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
  const long long t0 = h[0 * 16 + 3];
  const char* names[16] = {"S.issue.begin", "S.kfull.ok", "PV.vfull.ok", "PV.pfull.ok", "w2.wait.S", "w2.S.ok", "w2.max.ok",
                           "w2.exp.ok", "w2.P.done", "w6.wait.S", "w6.S.ok", "w6.max.ok", "w6.exp.ok", "w6.P.done", "K.load",
                           "V.load"};
  printf("%4s", "tile");
  for (int e = 0; e < 16; ++e) printf(" %13s", names[e]);
  printf("\n");
  for (int j = 0; j < 32; ++j) {
    printf("%4d", j + 16);
    for (int e = 0; e < 16; ++e) printf(" %13lld", h[j * 16 + e] ? h[j * 16 + e] - t0 : -1);
    printf("\n");
  }
  return 0;
}
