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
#define MD2_NOINLINE __device__ __noinline__
#else
#define MD2_DEVICE_BUILD 0
#define MD2_FN inline
#define MD2_HD inline
#define MD2_NOINLINE
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
// Keeps a computed address in a register pair: without it nvcc carries 64-bit element offsets and rebuilds the byte
// address (LEA + LEA.HI.X) for every load of a gather footprint.
MD2_FN const float* opaque_ptr(const float* p) {
  asm("" : "+l"(p));
  return p;
}
MD2_FN void atomic_add(float* p, float v) { atomicAdd(p, v); }
// one 16-byte reduction (red.global.add.v4.f32, sm_90+) for four consecutive floats; p is 16-byte aligned
MD2_FN void atomic_add4(float* p, const float* v) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
#else
MD2_FN float fadd(float a, float b) { return a + b; }
MD2_FN float fsub(float a, float b) { return a - b; }
MD2_FN float fmul(float a, float b) { return a * b; }
MD2_FN float ffma(float a, float b, float c) { return fmaf(a, b, c); }
MD2_FN float fdiv(float a, float b) { return a / b; }
MD2_FN float frcp(float a) { return 1.0f / a; }
MD2_FN float ld_ro(const float* p) { return *p; }
MD2_FN uint8_t ld_ro(const uint8_t* p) { return *p; }
MD2_FN const float* opaque_ptr(const float* p) { return p; }
MD2_FN void atomic_add(float* p, float v) { *p += v; }
MD2_FN void atomic_add4(float* p, const float* v) { for (int i = 0; i < 4; ++i) p[i] += v[i]; }
#endif

// ---- packed pairs: two independent fp32 lanes per instruction (add/mul/fma.rn.f32x2 on sm_100) ----
// Used to evaluate two source frames at once; every lane is rounded exactly like the scalar op.
#if MD2_DEVICE_BUILD
typedef float2 f2;
MD2_FN f2 mk2(float a, float b) { return make_float2(a, b); }
// Inline PTX with explicit .rn: nvcc contracts __fmul2_rn + __fadd2_rn into FFMA2 (observed in SASS),
// which would change the rounding sequence; PTX-level .rn operations are never fused.
MD2_FN unsigned long long f2_bits(f2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
MD2_FN f2 bits_f2(unsigned long long v) { return *reinterpret_cast<f2*>(&v); }
MD2_FN f2 fadd2(f2 a, f2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
// ptxas (12.9) fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry .rn (the scalar forms are
// never fused).  A product written as fma(a, b, +0) cannot be simplified to a mul (it differs for -0) and so
// cannot be fused with a following add; it rounds exactly like the mul (a -0 product becomes +0).
MD2_FN f2 fmul2(f2 a, f2 b) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(0ull));
  return bits_f2(d);
}
MD2_FN f2 ffma2(f2 a, f2 b, f2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
#else
struct f2 {
  float x, y;
};
MD2_FN f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
MD2_FN f2 fadd2(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
MD2_FN f2 fmul2(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
MD2_FN f2 ffma2(f2 a, f2 b, f2 c) { return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
MD2_FN f2 bc2(float a) { return mk2(a, a); }

// four consecutive floats with one 128-bit load (p is 16-byte aligned)
#if MD2_DEVICE_BUILD
typedef float4 f4;
#else
struct f4 {
  float x, y, z, w;
};
#endif
MD2_FN f4 ld4(const float* p) { return *reinterpret_cast<const f4*>(p); }
MD2_FN f2 fsub2(f2 a, f2 b) { return ffma2(b, bc2(-1.0f), a); }  // a - b, exact product, one rounding

MD2_HD int imin(int a, int b) { return a < b ? a : b; }
MD2_HD int imax(int a, int b) { return a > b ? a : b; }

}  // namespace md2
