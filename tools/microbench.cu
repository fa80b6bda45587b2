// FP32 issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (f32x2) vs FADD / FADD2 on sm_100a.
// Answers whether the fused loss (ALU-bound) should use packed fp32 math.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  float2 p0 = {a0, a1}, p1 = {a2, a3}, p2 = {a4, a5}, p3 = {a6, a7};
  const float2 m2 = {m, m}, c2 = {c, c};
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 8 scalar FFMA
      a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
      a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
    } else if (MODE == 1) {  // 4 FFMA2 = 8 fma
      p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
    } else if (MODE == 2) {  // 8 scalar FADD
      a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c);
      a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c);
    } else {  // 4 FADD2
      p0 = __fadd2_rn(p0, c2); p1 = __fadd2_rn(p1, c2); p2 = __fadd2_rn(p2, c2); p3 = __fadd2_rn(p3, c2);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y;
}

template <int MODE>
double run(float* d, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, iters);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return 148.0 * 8 * 256 * (double)iters * 8 / (ms * 1e-3) / 1e12;  // T lane-ops/s
}

int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int it = 20000;
  printf("{\"ffma_scalar_Tops\": %.2f, \"ffma2_packed_Tops\": %.2f, \"fadd_scalar_Tops\": %.2f, \"fadd2_packed_Tops\": %.2f}\n",
         run<0>(d, it), run<1>(d, it), run<2>(d, it), run<3>(d, it));
  return 0;
}
