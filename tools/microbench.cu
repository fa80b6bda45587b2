// FP32 issue-rate microbenchmark on sm_100a: scalar FFMA / FADD / FMUL vs the packed f32x2 forms, the cost of
// building a packed operand from a scalar (MOV), 32- vs 64-bit shared-memory loads, and the select / min-max ops
// the fused loss uses.  Answers which instruction mix the tile kernel (issue-bound) should prefer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
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
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
  return up(d);
}

// MODE: 0 FFMA x8 | 1 FFMA2 x4 | 2 FADD x8 | 3 FADD2 x4 | 4 FMUL x8 | 5 FMUL2 x4 | 6 FFMA2 x4 + 4 dependent-free MOV
//       7 LDS.32 x8 | 8 LDS.64 x4 | 9 FSEL x8 | 10 FMNMX x8 | 11 FFMA x4 + FADD x4 (fma pipe + ?) | 12 FFMA2 x4 + IADD x4
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ float2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(i * 1e-3f, 1.f);
  __syncthreads();
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  float2 p0 = {a0, a1}, p1 = {a2, a3}, p2 = {a4, a5}, p3 = {a6, a7};
  const float2 m2 = {m, m}, c2 = {c, c};
  int q0 = threadIdx.x, q1 = 1, q2 = 2, q3 = 3;
  const float* smf = reinterpret_cast<const float*>(sm);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
      a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
    } else if (MODE == 1) {
      p0 = fma2(p0, m2, c2); p1 = fma2(p1, m2, c2); p2 = fma2(p2, m2, c2); p3 = fma2(p3, m2, c2);
    } else if (MODE == 2) {
      a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c);
      a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c);
    } else if (MODE == 3) {
      p0 = add2(p0, c2); p1 = add2(p1, c2); p2 = add2(p2, c2); p3 = add2(p3, c2);
    } else if (MODE == 4) {
      a0 = __fmul_rn(a0, m); a1 = __fmul_rn(a1, m); a2 = __fmul_rn(a2, m); a3 = __fmul_rn(a3, m);
      a4 = __fmul_rn(a4, m); a5 = __fmul_rn(a5, m); a6 = __fmul_rn(a6, m); a7 = __fmul_rn(a7, m);
    } else if (MODE == 5) {
      p0 = mul2(p0, m2); p1 = mul2(p1, m2); p2 = mul2(p2, m2); p3 = mul2(p3, m2);
    } else if (MODE == 6) {
      // packed operand built from a scalar every time: (a, a) needs a MOV into the odd register
      float2 b0 = {p0.x, p0.x}, b1 = {p1.x, p1.x}, b2 = {p2.x, p2.x}, b3 = {p3.x, p3.x};
      p0 = fma2(b0, m2, p0); p1 = fma2(b1, m2, p1); p2 = fma2(b2, m2, p2); p3 = fma2(b3, m2, p3);
    } else if (MODE == 7) {
      a0 += smf[(q0 + 0) & 2047]; a1 += smf[(q0 + 32) & 2047]; a2 += smf[(q0 + 64) & 2047]; a3 += smf[(q0 + 96) & 2047];
      a4 += smf[(q0 + 128) & 2047]; a5 += smf[(q0 + 160) & 2047]; a6 += smf[(q0 + 192) & 2047]; a7 += smf[(q0 + 224) & 2047];
      q0 += 256;
    } else if (MODE == 8) {
      p0 = add2(p0, sm[(q0 + 0) & 1023]); p1 = add2(p1, sm[(q0 + 32) & 1023]); p2 = add2(p2, sm[(q0 + 64) & 1023]);
      p3 = add2(p3, sm[(q0 + 96) & 1023]);
      q0 += 128;
    } else if (MODE == 9) {
      a0 = a0 > a1 ? a2 : a0; a1 = a1 > a2 ? a3 : a1; a2 = a2 > a3 ? a4 : a2; a3 = a3 > a4 ? a5 : a3;
      a4 = a4 > a5 ? a6 : a4; a5 = a5 > a6 ? a7 : a5; a6 = a6 > a7 ? a0 : a6; a7 = a7 > a0 ? a1 : a7;
    } else if (MODE == 10) {
      a0 = fminf(a0, a1); a1 = fmaxf(a1, a2); a2 = fminf(a2, a3); a3 = fmaxf(a3, a4);
      a4 = fminf(a4, a5); a5 = fmaxf(a5, a6); a6 = fminf(a6, a7); a7 = fmaxf(a7, a0);
    } else if (MODE == 11) {
      a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
      a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c);
    } else if (MODE == 12) {
      p0 = fma2(p0, m2, c2); p1 = fma2(p1, m2, c2); p2 = fma2(p2, m2, c2); p3 = fma2(p3, m2, c2);
      q0 = (q0 + q1) ^ q2; q1 = (q1 + q2) ^ q3; q2 = (q2 + q3) ^ q0; q3 = (q3 + q0) ^ q1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + q0 + q1 + q2 + q3;
}

template <int MODE>
double run(float* d, int iters, int warps_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int nt = 32 * warps_per_sm;
  k<MODE><<<148, nt>>>(d, iters);
  cudaEventRecord(e0);
  k<MODE><<<148, nt>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-instructions per cycle per SM at 1.965 GHz; 8 "lane-ops" per iteration in every mode
  return 148.0 * nt * (double)iters * 8 / (ms * 1e-3) / 1e12;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int it = 20000;
  const char* names[13] = {"ffma", "ffma2", "fadd", "fadd2", "fmul", "fmul2", "ffma2_plus_mov", "lds32", "lds64_as_2",
                           "fsel", "fmnmx", "ffma_fadd_mix", "ffma2_plus_int"};
  for (int w = 8; w <= 32; w *= 2) {
    double r[13] = {run<0>(d, it, w), run<1>(d, it, w), run<2>(d, it, w), run<3>(d, it, w), run<4>(d, it, w),
                    run<5>(d, it, w), run<6>(d, it, w), run<7>(d, it, w), run<8>(d, it, w), run<9>(d, it, w),
                    run<10>(d, it, w), run<11>(d, it, w), run<12>(d, it, w)};
    printf("{\"warps_per_sm\": %d", w);
    for (int i = 0; i < 13; ++i) printf(", \"%s_Tlaneops\": %.2f", names[i], r[i]);
    printf("}\n");
  }
  return 0;
}
