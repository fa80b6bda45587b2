// Dependent-issue latency of the fp32 ops the tile kernel chains (scalar vs packed f32x2), one warp per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_lat tools/microbench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ float2 up(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return up(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
  return up(d);
}
__device__ __forceinline__ float fadd_v(float a, float b) {
  float d;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float ffma_v(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// MODE 0: FADD chain  1: FADD2 chain  2: FFMA chain  3: FFMA2 chain  4: 4 independent FADD2 chains  5: 4 indep FADD chains
// 6: LDS dependent (pointer chase)  7: MUFU.RCP chain  8: FSEL/FSETP chain  9: 2 independent FADD2 chains
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ int chase[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase[i] = (i * 33 + 32) & 1023;
  __syncthreads();
  float a = threadIdx.x * 1e-3f + 1.f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  float2 p = {a, b}, q = {c, d}, r = {a, c}, s = {b, d};
  const float2 m2 = {0.999f, 0.999f}, c2 = {1e-3f, 1e-3f};
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) a = fadd_v(a, 1e-3f);
      if (MODE == 1) p = add2(p, c2);
      if (MODE == 2) a = ffma_v(a, 0.999f, 1e-3f);
      if (MODE == 3) p = fma2(p, m2, c2);
      if (MODE == 4) { p = add2(p, c2); q = add2(q, c2); r = add2(r, c2); s = add2(s, c2); }
      if (MODE == 5) { a = fadd_v(a, 1e-3f); b = fadd_v(b, 1e-3f); c = fadd_v(c, 1e-3f); d = fadd_v(d, 1e-3f); }
      if (MODE == 6) idx = chase[idx];
      if (MODE == 7) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
      if (MODE == 8) a = a > b ? c : a + 0.f;
      if (MODE == 9) { p = add2(p, c2); q = add2(q, c2); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + p.x + p.y + q.x + q.y + r.x + r.y + s.x + s.y + idx;
}
template <int MODE>
double run(float* d, long long* c, int nper) {
  const int iters = 2000;
  k<MODE><<<1, 32>>>(d, c, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  return (double)h / (iters * 16.0 * nper);
}
int main() {
  float* d; long long* c;
  cudaMalloc(&d, 4096 * 4); cudaMalloc(&c, 64);
  printf("{\"cycles_per_instr_one_warp\": {\"fadd_chain\": %.2f, \"fadd2_chain\": %.2f, \"ffma_chain\": %.2f, \"ffma2_chain\": %.2f, "
         "\"fadd2_x4_indep\": %.2f, \"fadd_x4_indep\": %.2f, \"lds_chase\": %.2f, \"mufu_rcp_chain\": %.2f, \"fsel_chain\": %.2f, \"fadd2_x2_indep\": %.2f}}\n",
         run<0>(d, c, 1), run<1>(d, c, 1), run<2>(d, c, 1), run<3>(d, c, 1), run<4>(d, c, 4), run<5>(d, c, 4), run<6>(d, c, 1),
         run<7>(d, c, 1), run<8>(d, c, 1), run<9>(d, c, 2));
  return 0;
}
