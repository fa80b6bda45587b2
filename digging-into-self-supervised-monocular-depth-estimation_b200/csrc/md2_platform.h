// md2_platform.h - one source, two compilations.
//
// The tile code in md2_tile.cuh is written once.  nvcc compiles it for sm_100a (the
// product).  tests/host_emu compiles the SAME phase functions with g++ and runs the
// threads of a CTA as a loop between barriers, so that the CPU-only test-suite can
// check the kernel logic against the oracle without a GPU.  The host build is test
// infrastructure only: the shipped library contains no CPU path.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MD2_DEVICE_BUILD 1
#define MD2_FN __device__ __forceinline__
#define MD2_HD __host__ __device__ __forceinline__
#else
#define MD2_DEVICE_BUILD 0
#define MD2_FN inline
#define MD2_HD inline
#endif

namespace md2 {

// ---- correctly rounded fp32 primitives with no FMA contraction --------------------
// The reference evaluates the SSIM / projection chains as separate ATen kernels, i.e.
// every operation is rounded to fp32 on its own.  These wrappers pin that rounding
// sequence; the host build is compiled with -ffp-contract=off.
#if MD2_DEVICE_BUILD
MD2_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
MD2_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
MD2_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
MD2_FN float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
MD2_FN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
MD2_FN float frcp(float a) { return __frcp_rn(a); }
MD2_FN float ld_ro(const float* p) { return __ldg(p); }
MD2_FN uint8_t ld_ro(const uint8_t* p) { return __ldg(p); }
MD2_FN void atomic_add(float* p, float v) { atomicAdd(p, v); }
#else
MD2_FN float fadd(float a, float b) { return a + b; }
MD2_FN float fsub(float a, float b) { return a - b; }
MD2_FN float fmul(float a, float b) { return a * b; }
MD2_FN float ffma(float a, float b, float c) { return fmaf(a, b, c); }
MD2_FN float fdiv(float a, float b) { return a / b; }
MD2_FN float frcp(float a) { return 1.0f / a; }
MD2_FN float ld_ro(const float* p) { return *p; }
MD2_FN uint8_t ld_ro(const uint8_t* p) { return *p; }
MD2_FN void atomic_add(float* p, float v) { *p += v; }
#endif

MD2_HD int imin(int a, int b) { return a < b ? a : b; }
MD2_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- counter-based N(0,1) for the auto-mask tie-breaker when no noise is supplied ----
MD2_FN uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
MD2_FN float gauss_from_counter(uint64_t seed, uint64_t counter) {
  const uint64_t h = splitmix64(seed ^ splitmix64(counter));
  const float u1 = ((float)((h >> 40) & 0xFFFFFF) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  const float u2 = (float)((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);            // [0,1)
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

}  // namespace md2
