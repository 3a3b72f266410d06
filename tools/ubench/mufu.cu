// Micro-benchmark: exponentials per clock per SM sub-partition for the softmax inner loop variants.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/mufu tools/ubench/mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%5, %5};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{\n.reg .b64 ra, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rd, {%0, %1};\nadd.rn.f32x2 rd, rd, ra;\nmov.b64 {%0, %1}, rd;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void fadd2o(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// packed polynomial 2^x for a pair
__device__ __forceinline__ void ex2_poly2(float& p0, float& p1, float x0, float x1) {
  x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
  float t0, t1, u0, u1, f0, f1;
  fadd2o(t0, t1, x0, x1, 12582912.f, 12582912.f);
  fadd2o(u0, u1, t0, t1, -12582912.f, -12582912.f);
  fadd2o(f0, f1, x0, x1, -u0, -u1);
  float q0, q1;
  ffma2v(q0, q1, f0, f1, 0.05520550534f, 0.05520550534f, 0.24261397123f, 0.24261397123f);
  ffma2v(q0, q1, q0, q1, f0, f1, 0.69325476885f, 0.69325476885f);
  ffma2v(q0, q1, q0, q1, f0, f1, 0.99992769957f, 0.99992769957f);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// MODE 0: f32 MUFU, 1: f16x2 MUFU, 2: bf16x2 MUFU, 3: 1/4 poly (packed), 4: 1/2 poly, 5: 1/8 poly, 6: f16x2 + f32 row sum
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* in, uint32_t* out, long long* cyc, int iters, float c, float nm) {
  float r[64];
  for (int i = 0; i < 64; ++i) r[i] = in[(threadIdx.x * 64 + i) & 1023];
  uint32_t acc[32];
  for (int i = 0; i < 32; ++i) acc[i] = 0;
  float l0 = 0.f, l1 = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float x0, x1;
      ffma2(x0, x1, r[i], r[i + 1], c, nm);
      if (MODE == 0 || MODE >= 3 && MODE <= 5) {
        float p0, p1;
        const int e = i & 15;
        const bool poly = (MODE == 3 && (e == 4 || e == 12)) || (MODE == 4 && (e == 2 || e == 6 || e == 10 || e == 14)) ||
                          (MODE == 5 && e == 6);
        if (poly) ex2_poly2(p0, p1, x0, x1);
        else { p0 = ex2(x0); p1 = ex2(x1); }
        fadd2(l0, l1, p0, p1);
        __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
        acc[i >> 1] ^= *(uint32_t*)&hh;
      } else if (MODE == 1 || MODE == 6) {
        __half2 hx = __floats2half2_rn(x0, x1);
        uint32_t pp = ex2_h2(*(uint32_t*)&hx);
        acc[i >> 1] ^= pp;
        if (MODE == 6) {
          float2 pf = __half22float2(*(__half2*)&pp);
          fadd2(l0, l1, pf.x, pf.y);
        }
      } else {
        __nv_bfloat162 hx = __floats2bfloat162_rn(x0, x1);
        uint32_t pp = ex2_b2(*(uint32_t*)&hx);
        acc[i >> 1] ^= pp;
      }
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] += 1e-3f;       // keep the loop body live
  }
  long long t1 = clock64();
  uint32_t s = __float_as_uint(l0 + l1);
  for (int i = 0; i < 32; ++i) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(in, out, cyc, iters, 0.09f, -3.f);
  k<MODE><<<148, warps * 32>>>(in, out, cyc, iters, 0.09f, -3.f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // elements per SMSP = warps/4 * 32 lanes * 64 * iters
  const double el = (double)warps / 4 * 32 * 64 * iters;
  printf("%-28s warps/SM %2d: %.3f elements/clk/SMSP  (%.1f clk per 128x128 tile per SMSP)  err=%s\n", name, warps, el / c,
         4096.0 / (el / c), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  for (int w : {4, 8}) {
    run<0>("f32 MUFU", w);
    run<1>("f16x2 MUFU", w);
    run<6>("f16x2 MUFU + f32 sum", w);
    run<2>("bf16x2 MUFU", w);
    run<5>("f32 MUFU, 1/8 poly", w);
    run<3>("f32 MUFU, 1/4 poly", w);
    run<4>("f32 MUFU, 1/2 poly", w);
  }
  return 0;
}
