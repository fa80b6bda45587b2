// md2_pipeline.cu - on-device colour pyramid: flip + Pillow-exact antialiased resize of the
// decoded uint8 image to every pyramid level + ToTensor, replacing kitti_mono.py:283-288,296-304,347-362.
//
// Host part: the Lanczos coefficient tables of Pillow's resampler (double precision, libm), one per axis and
// level.  Device part, all integer / byte work:
//   to_planar   [N,Hin,Win,3] u8 -> [N*3*Hin][pitch] u8 planes (flip applied here), so that every later pass
//               sees independent rows whose taps are contiguous bytes
//   pyramid_h   per level: horizontal pass with dp4a.u32.s32 - four taps per instruction on aligned 32-bit words;
//               the 23-bit fixed-point coefficients are split into three signed 8-bit digits
//               (c = d0 + 2^8 d1 + 2^16 d2), one dp4a accumulator per digit, recombined exactly in int32
//   pyramid_v   per level: vertical pass on four byte columns per thread (one 32-bit load per tap row),
//               + ToTensor (/255) straight into the CHW float output (one 16-byte store)
// Results are bit-identical to Pillow (tests/test_gpu_pipeline.py).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/md2_pipeline.h"
#include "md2_host.h"
#include "md2_nvtx.h"

namespace md2 {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS
constexpr int kRowsPerThread = 8;           // rows that share one set of coefficient loads in pyramid_h

// ---- table layout (ints), per level:
//   X axis: [w0: out] first aligned input word of each output pixel
//           [cd: 3 digits][kw words][out]   packed signed 8-bit coefficient digits per aligned input word
//   Y axis: [bounds: 2 * out][coeffs: out * ksize]   as in Pillow (scalar fallback)
//           [cd: out][ng groups of 4 tap rows][3 digits]   packed signed 8-bit digits of 4 consecutive taps
struct LevelTab {
  int wout, hout, ksx, ksy, kw, ng;
  size_t x_w0, x_cd, y_bounds, y_coef, y_cd;  // offsets in ints
};

static inline int lanczos_ksize(int in, int out) {
  double fs = (double)in / (double)out;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(3.0 * fs) * 2 + 1;
}

MD2_HD int planar_pitch(int Win) { return ((Win + 3) / 4) * 4 + 4; }  // multiple of 4, >= Win + 4 (zero padded)

static inline size_t pyramid_layout(const md2_pyramid_cfg& c, LevelTab tab[kMaxScales]) {
  size_t off = 0;
  for (int s = 0; s < c.scales; ++s) {
    LevelTab t;
    t.wout = c.W >> s;
    t.hout = c.H >> s;
    t.ksx = lanczos_ksize(c.Win, t.wout);
    t.ksy = lanczos_ksize(c.Hin, t.hout);
    t.kw = (t.ksx + 3) / 4 + 1;  // aligned words that can hold ksx consecutive bytes at any alignment
    t.x_w0 = off;
    off += t.wout;
    t.x_cd = off;
    off += (size_t)3 * t.kw * t.wout;
    t.y_bounds = off;
    off += 2 * (size_t)t.hout;
    t.y_coef = off;
    off += (size_t)t.hout * t.ksy;
    t.ng = (t.ksy + 3) / 4;
    t.y_cd = off;
    off += (size_t)t.hout * t.ng * 3;
    if (tab) tab[s] = t;
  }
  return off;
}

static inline int validate_pyramid(const md2_pyramid_cfg* c) {
  if (!c) return MD2_ERR_NULL;
  if (c->N < 1 || c->Hin < 1 || c->Win < 1 || c->H < 1 || c->W < 1) return MD2_ERR_SHAPE;
  if (c->scales < 1 || c->scales > kMaxScales) return MD2_ERR_SHAPE;
  if ((c->H >> (c->scales - 1)) < 1 || (c->W >> (c->scales - 1)) < 1) return MD2_ERR_SHAPE;
  if ((long long)c->N * 3 * c->Hin * planar_pitch(c->Win) > 0x7fffffffLL) return MD2_ERR_SHAPE;
  if ((long long)c->N * 3 * c->Hin > 65535LL * kRowsPerThread || c->Hin > 65535 || c->N > 21845) return MD2_ERR_SHAPE;  // grid.y / .z
  if (lanczos_ksize(c->Win, c->W >> (c->scales - 1)) > 1024 || lanczos_ksize(c->Hin, c->H >> (c->scales - 1)) > 1024)
    return MD2_ERR_SHAPE;
  return 0;
}

// Pillow's lanczos_filter / sinc_filter
static inline double lanczos3(double x) {
  if (-3.0 <= x && x < 3.0) {
    if (x == 0.0) return 1.0;
    const double a = x * M_PI, b = a / 3.0;
    return (sin(a) / a) * (sin(b) / b);
  }
  return 0.0;
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for one output position (box = the whole image)
static void axis_coeffs(int in, int out, int xx, int ksize, int* k, int& xmin, int& cnt) {
  const double scale = (double)in / (double)out;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * fs, ss = 1.0 / fs;
  const double center = (xx + 0.5) * scale;
  double w[1024];
  xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in) xmax = in;
  cnt = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < cnt; ++x) {
    w[x] = lanczos3((x + xmin - center + 0.5) * ss);
    ww += w[x];
  }
  for (int x = 0; x < ksize; ++x) {
    double v = x < cnt ? w[x] : 0.0;
    if (x < cnt && ww != 0.0) v /= ww;
    k[x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
  }
}

// three signed 8-bit digits of four coefficients, packed little-endian: c = d0 + 2^8 d1 + 2^16 d2
static void pack_digits(const int cf[4], unsigned packed[3]) {
  packed[0] = packed[1] = packed[2] = 0u;
  for (int b = 0; b < 4; ++b) {
    const int d0 = ((cf[b] + 128) & 255) - 128;
    const int c1 = (cf[b] - d0) >> 8;
    const int d1 = ((c1 + 128) & 255) - 128;
    const int d2 = (c1 - d1) >> 8;  // |cf| < 2^23  =>  |d2| <= 64
    packed[0] |= (unsigned)(d0 & 255) << (8 * b);
    packed[1] |= (unsigned)(d1 & 255) << (8 * b);
    packed[2] |= (unsigned)(d2 & 255) << (8 * b);
  }
}

static void fill_level(const md2_pyramid_cfg& c, const LevelTab& t, int* base) {
  int k[1024 + 4];
  // Y axis: Pillow's layout
  for (int y = 0; y < t.hout; ++y) {
    int ymin, cnt;
    axis_coeffs(c.Hin, t.hout, y, t.ksy, k, ymin, cnt);
    base[t.y_bounds + 2 * y] = ymin;
    base[t.y_bounds + 2 * y + 1] = cnt;
    for (int j = 0; j < t.ksy; ++j) base[t.y_coef + (size_t)y * t.ksy + j] = k[j];
    for (int g = 0; g < t.ng; ++g) {
      int cf[4];
      unsigned packed[3];
      for (int b = 0; b < 4; ++b) cf[b] = (4 * g + b < cnt) ? k[4 * g + b] : 0;
      pack_digits(cf, packed);
      for (int d = 0; d < 3; ++d) base[t.y_cd + ((size_t)y * t.ng + g) * 3 + d] = (int)packed[d];
    }
  }
  // X axis: per aligned input word, three packed signed-digit words
  for (int x = 0; x < t.wout; ++x) {
    int xmin, cnt;
    axis_coeffs(c.Win, t.wout, x, t.ksx, k, xmin, cnt);
    const int w0 = xmin >> 2;
    base[t.x_w0 + x] = w0;
    for (int w = 0; w < t.kw; ++w) {
      unsigned packed[3];
      int cf[4];
      for (int b = 0; b < 4; ++b) {
        const int j = 4 * (w0 + w) + b - xmin;
        cf[b] = (j >= 0 && j < cnt) ? k[j] : 0;
      }
      pack_digits(cf, packed);
      for (int d = 0; d < 3; ++d) base[t.x_cd + ((size_t)d * t.kw + w) * t.wout + x] = (int)packed[d];
    }
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// v / 255 correctly rounded for the 256 possible byte values (transforms.ToTensor's .div(255)): one Newton
// correction of v * fl(1/255); equal to IEEE division for every v in [0, 255] (tests/test_gpu_pipeline.py compares
// every output value with Pillow + ToTensor)
__device__ __forceinline__ float div255(int v) {
  const float r = 1.0f / 255.0f, x = (float)v;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-255.0f, q, x), r, q);
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// All index arithmetic below is 32-bit (validate_pyramid bounds every extent by 2^31) and the row / plane
// coordinates come from blockIdx.y / .z: 64-bit divisions were 2/3 of the instructions of the first version.

// [N,Hin,Win,3] u8 -> planar rows [(n*3 + c)*Hin + y][pitch], flipped if asked, zero padded.
// grid (words / 128, Hin, N); a thread turns 4 RGB pixels (12 bytes) into one word of each plane.
__global__ void __launch_bounds__(128) pyramid_to_planar(int Hin, int Win, int pitch, const uint8_t* __restrict__ img,
                                                         const uint8_t* __restrict__ flip, unsigned* planar) {
  const int pw = pitch >> 2;
  const int w = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (w >= pw) return;
  const uint8_t* src = img + ((size_t)n * Hin + y) * Win * 3;
  const bool fl = flip && flip[n];
  unsigned v[3] = {0u, 0u, 0u};
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int x = 4 * w + b;
    if (x < Win) {
      const uint8_t* px = src + (fl ? Win - 1 - x : x) * 3;
      v[0] |= (unsigned)px[0] << (8 * b);
      v[1] |= (unsigned)px[1] << (8 * b);
      v[2] |= (unsigned)px[2] << (8 * b);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) planar[((size_t)(n * 3 + c) * Hin + y) * pw + w] = v[c];
}

// horizontal pass on planar rows: [rows][pitch] u8 -> [rows][wout] u8.  grid (wout / 128, row groups); a thread owns
// one x_out of kRowsPerThread consecutive rows, so that its three coefficient words per input word are loaded once
template <int KW>
__global__ void __launch_bounds__(128) pyramid_h(int rows, int pitch, int wout, int kw_rt, const unsigned* __restrict__ planar,
                                                 const int* __restrict__ w0_tab, const int* __restrict__ cd, uint8_t* tmp) {
  const int kw = KW > 0 ? KW : kw_rt;
  const int pw = pitch >> 2;
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= wout) return;
  const int r0 = blockIdx.y * kRowsPerThread;
  int a0[kRowsPerThread], a1[kRowsPerThread], a2[kRowsPerThread];
  const unsigned* src[kRowsPerThread];
  const int w0 = w0_tab[x];
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) {
    a0[r] = a1[r] = a2[r] = 0;
    src[r] = planar + (size_t)imin(r0 + r, rows - 1) * pw + w0;
  }
  const int* c0p = cd + x;
  const int* c1p = c0p + kw * wout;
  const int* c2p = c1p + kw * wout;
#pragma unroll
  for (int w = 0; w < (KW > 0 ? KW : 1); ++w) {
    for (int ww = w; ww < kw; ww += (KW > 0 ? kw : 1)) {  // compile-time trip count when KW > 0, run-time loop otherwise
      const int c0 = c0p[ww * wout], c1 = c1p[ww * wout], c2 = c2p[ww * wout];
#pragma unroll
      for (int r = 0; r < kRowsPerThread; ++r) {
        const unsigned px = src[r][ww];
        a0[r] = dp4a_us(px, c0, a0[r]);
        a1[r] = dp4a_us(px, c1, a1[r]);
        a2[r] = dp4a_us(px, c2, a2[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r)
    if (r0 + r < rows) {
      // exact modulo 2^32, and the true sum fits int32 (Pillow relies on the same bound)
      const unsigned ss = (1u << (kPrecisionBits - 1)) + (unsigned)a0[r] + (unsigned)a1[r] * 256u + (unsigned)a2[r] * 65536u;
      tmp[(size_t)(r0 + r) * wout + x] = (uint8_t)clip8((int)ss);
    }
}

// vertical pass + ToTensor on planar rows: [(n*3+c)*Hin + y][wout] u8 -> [N,3,hout,wout] f32; VEC byte columns per
// thread.  grid (hout * (wout / VEC) / 128, planes)
template <int VEC>
__global__ void __launch_bounds__(128) pyramid_v(int Hin, int hout, int wout, int ksize, const uint8_t* __restrict__ tmp,
                                                 const int* __restrict__ bounds, const int* __restrict__ kk, float* out) {
  const int wv = wout / VEC;
  const int i = blockIdx.x * 128 + threadIdx.x, pl = blockIdx.y;
  if (i >= hout * wv) return;
  const int y = i / wv, xv = i - y * wv;
  const int ymin = bounds[2 * y], cnt = bounds[2 * y + 1];
  const int* k = kk + y * ksize;
  const uint8_t* src = tmp + ((size_t)pl * Hin + ymin) * wout + xv * VEC;
  int acc[VEC];
#pragma unroll
  for (int b = 0; b < VEC; ++b) acc[b] = 1 << (kPrecisionBits - 1);
  for (int j = 0; j < cnt; ++j) {
    const int c = k[j];
    if (VEC == 4) {
      const unsigned px = *reinterpret_cast<const unsigned*>(src);
      acc[0] += (int)__byte_perm(px, 0, 0x4440) * c;
      acc[1] += (int)__byte_perm(px, 0, 0x4441) * c;
      acc[2] += (int)__byte_perm(px, 0, 0x4442) * c;
      acc[3] += (int)__byte_perm(px, 0, 0x4443) * c;
    } else {
      acc[0] += src[0] * c;
    }
    src += wout;
  }
  float* o = out + ((size_t)pl * hout + y) * wout + xv * VEC;  // transforms.ToTensor: uint8 -> float32, .div(255)
  if (VEC == 4) {
    *reinterpret_cast<float4*>(o) = make_float4(div255(clip8(acc[0])), div255(clip8(acc[1])), div255(clip8(acc[2])), div255(clip8(acc[3])));
  } else {
    o[0] = div255(clip8(acc[0]));
  }
}

// The same vertical pass with dp4a: four tap rows of four byte columns are transposed in registers (8 PRMT) so
// that each column's four taps sit in one word, then 3 digit dp4a per column.  Rows past the window carry zero
// coefficients (they may belong to the next plane or to the slack behind the buffer).
__global__ void __launch_bounds__(128) pyramid_v_dp4a(int Hin, int hout, int wout, int ng, const uint8_t* __restrict__ tmp,
                                                      const int* __restrict__ bounds, const int* __restrict__ cd, float* out) {
  const int wv = wout >> 2;
  const int i = blockIdx.x * 128 + threadIdx.x, pl = blockIdx.y;
  if (i >= hout * wv) return;
  const int y = i / wv, xv = i - y * wv;
  const int ymin = bounds[2 * y];
  const int* c = cd + (size_t)y * ng * 3;
  const unsigned* src = reinterpret_cast<const unsigned*>(tmp + ((size_t)pl * Hin + ymin) * wout) + xv;
  int a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  for (int g = 0; g < ng; ++g) {
    const unsigned r0 = src[0], r1 = src[wv], r2 = src[2 * wv], r3 = src[3 * wv];
    const int c0 = c[0], c1 = c[1], c2 = c[2];
    const unsigned lo01 = __byte_perm(r0, r1, 0x5140), hi01 = __byte_perm(r0, r1, 0x7362);
    const unsigned lo23 = __byte_perm(r2, r3, 0x5140), hi23 = __byte_perm(r2, r3, 0x7362);
    const unsigned t[4] = {__byte_perm(lo01, lo23, 0x5410), __byte_perm(lo01, lo23, 0x7632), __byte_perm(hi01, hi23, 0x5410),
                           __byte_perm(hi01, hi23, 0x7632)};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      a0[b] = dp4a_us(t[b], c0, a0[b]);
      a1[b] = dp4a_us(t[b], c1, a1[b]);
      a2[b] = dp4a_us(t[b], c2, a2[b]);
    }
    src += 4 * wv;
    c += 3;
  }
  float v[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const unsigned ss = (1u << (kPrecisionBits - 1)) + (unsigned)a0[b] + (unsigned)a1[b] * 256u + (unsigned)a2[b] * 65536u;
    v[b] = div255(clip8((int)ss));
  }
  *reinterpret_cast<float4*>(out + ((size_t)pl * hout + y) * wout + xv * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

template <int KW>
static void launch_h(int rows, int pitch, const LevelTab& t, const unsigned* planar, const int* tables, uint8_t* tmp, cudaStream_t st) {
  const dim3 grid((t.wout + 127) / 128, (rows + kRowsPerThread - 1) / kRowsPerThread);
  pyramid_h<KW><<<grid, 128, 0, st>>>(rows, pitch, t.wout, t.kw, planar, tables + t.x_w0, tables + t.x_cd, tmp);
}


// ---- ToTensor on the device --------------------------------------------------------------------------------------
// transforms.ToTensor() of kitti_mono.py:283 (called at :352, :364): uint8 HWC -> float32 CHW, .div(255).  The loader
// keeps the resized PIL images as bytes and the step uploads those (a quarter of the float bytes); one launch converts
// every image group of the batch.  Four pixels per thread where the row length allows it: three aligned 32-bit loads
// (12 bytes = 4 RGB pixels), one 16-byte store per channel plane.
struct ToTensorJobs {
  const uint8_t* src[MD2_TO_TENSOR_MAX];
  float* dst[MD2_TO_TENSOR_MAX];
  int hw[MD2_TO_TENSOR_MAX];          // pixels per image
  int n[MD2_TO_TENSOR_MAX];           // images
  unsigned first_block[MD2_TO_TENSOR_MAX + 1];
  int count;
};

__global__ void __launch_bounds__(256) to_tensor_kernel(const __grid_constant__ ToTensorJobs jobs) {
  int j = 0;
  while (j + 1 < jobs.count && blockIdx.x >= jobs.first_block[j + 1]) ++j;
  const int hw = jobs.hw[j];
  const size_t total = (size_t)jobs.n[j] * hw;  // pixels of the group
  const uint8_t* __restrict__ src = jobs.src[j];
  float* __restrict__ dst = jobs.dst[j];
  const size_t q = ((size_t)(blockIdx.x - jobs.first_block[j]) * 256 + threadIdx.x) * 4;  // first pixel of this thread
  if (q >= total) return;
  const size_t img = q / hw;
  const int p = (int)(q - img * hw);
  float* o = dst + img * 3 * (size_t)hw + p;
  if ((hw & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src + q * 3);
    const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    *reinterpret_cast<float4*>(o) =
        make_float4(div255(a & 255), div255(a >> 24), div255((b >> 16) & 255), div255((c >> 8) & 255));
    *reinterpret_cast<float4*>(o + hw) =
        make_float4(div255((a >> 8) & 255), div255(b & 255), div255(b >> 24), div255((c >> 16) & 255));
    *reinterpret_cast<float4*>(o + 2 * (size_t)hw) =
        make_float4(div255((a >> 16) & 255), div255((b >> 8) & 255), div255(c & 255), div255(c >> 24));
  } else {
    for (int k = 0; k < 4 && q + k < total; ++k) {
      const size_t qq = q + k;
      const size_t im = qq / hw;
      const int pp = (int)(qq - im * hw);
      for (int ch = 0; ch < 3; ++ch) dst[(im * 3 + ch) * (size_t)hw + pp] = div255(src[qq * 3 + ch]);
    }
  }
}

}  // namespace md2

using namespace md2;

extern "C" {

size_t md2_pyramid_tables_bytes(const md2_pyramid_cfg* cfg) {
  if (validate_pyramid(cfg) != 0) return 0;
  return pyramid_layout(*cfg, nullptr) * sizeof(int);
}

int md2_pyramid_tables_fill(const md2_pyramid_cfg* cfg, void* host_tables) {
  const int v = validate_pyramid(cfg);
  if (v != 0) return v;
  if (!host_tables) return MD2_ERR_NULL;
  LevelTab tab[kMaxScales];
  pyramid_layout(*cfg, tab);
  for (int s = 0; s < cfg->scales; ++s) fill_level(*cfg, tab[s], (int*)host_tables);
  return 0;
}

size_t md2_pyramid_workspace_bytes(const md2_pyramid_cfg* cfg) {
  if (validate_pyramid(cfg) != 0) return 0;
  const size_t rows = (size_t)cfg->N * 3 * cfg->Hin;
  // planar copy of the input (+ one word of slack) and the level-0 intermediate, both 256-byte aligned
  const size_t planar = ((rows * planar_pitch(cfg->Win) + 4 * 1024 + 255) / 256) * 256;
  // level-0 intermediate + slack rows that the dp4a vertical pass may touch behind the last plane (zero coefficients)
  const size_t slack = (size_t)(lanczos_ksize(cfg->Hin, cfg->H >> (cfg->scales - 1)) + 8) * cfg->W;
  return planar + ((rows * cfg->W + slack + 255) / 256) * 256;
}

int md2_color_pyramid(const md2_pyramid_cfg* cfg, const uint8_t* images, const uint8_t* flip, const void* device_tables,
                      float* const* out, void* workspace, md2_stream_t stream) {
  const md2::NvtxRange range("md2_color_pyramid");
  const int v = validate_pyramid(cfg);
  if (v != 0) return v;
  if (!images || !device_tables || !out) return MD2_ERR_NULL;
  if (!workspace) return MD2_ERR_WORKSPACE;
  for (int s = 0; s < cfg->scales; ++s)
    if (!out[s]) return MD2_ERR_NULL;
  LevelTab tab[kMaxScales];
  pyramid_layout(*cfg, tab);
  cudaStream_t st = (cudaStream_t)stream;
  const int* tables = (const int*)device_tables;
  const int rows = cfg->N * 3 * cfg->Hin, pitch = planar_pitch(cfg->Win);
  unsigned* planar = (unsigned*)workspace;
  const size_t planar_bytes = (((size_t)rows * pitch + 4 * 1024 + 255) / 256) * 256;
  uint8_t* tmp = (uint8_t*)workspace + planar_bytes;
  // the slack after the last planar row is read (times zero coefficients) by the last rows' trailing words
  cudaMemsetAsync((uint8_t*)workspace + (size_t)rows * pitch, 0, planar_bytes - (size_t)rows * pitch, st);
  pyramid_to_planar<<<dim3(((pitch >> 2) + 127) / 128, cfg->Hin, cfg->N), 128, 0, st>>>(cfg->Hin, cfg->Win, pitch, images, flip, planar);
  for (int s = 0; s < cfg->scales; ++s) {
    const LevelTab& t = tab[s];
    switch (t.kw) {  // the KITTI 375x1242 -> 192x640 pyramid has 5 / 8 / 14 / 25 words per output pixel
      case 5: launch_h<5>(rows, pitch, t, planar, tables, tmp, st); break;
      case 8: launch_h<8>(rows, pitch, t, planar, tables, tmp, st); break;
      case 14: launch_h<14>(rows, pitch, t, planar, tables, tmp, st); break;
      case 25: launch_h<25>(rows, pitch, t, planar, tables, tmp, st); break;
      default: launch_h<0>(rows, pitch, t, planar, tables, tmp, st); break;
    }
    if (t.wout % 4 == 0)
      pyramid_v_dp4a<<<dim3((t.hout * (t.wout / 4) + 127) / 128, cfg->N * 3), 128, 0, st>>>(
          cfg->Hin, t.hout, t.wout, t.ng, tmp, tables + t.y_bounds, tables + t.y_cd, out[s]);
    else
      pyramid_v<1><<<dim3((t.hout * t.wout + 127) / 128, cfg->N * 3), 128, 0, st>>>(
          cfg->Hin, t.hout, t.wout, t.ksy, tmp, tables + t.y_bounds, tables + t.y_coef, out[s]);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int md2_to_tensor(int count, const md2_u8_images* groups, md2_stream_t stream) {
  const md2::NvtxRange range("md2_to_tensor");
  if (count == 0) return 0;
  if (!groups) return MD2_ERR_NULL;
  if (count < 0 || count > MD2_TO_TENSOR_MAX) return MD2_ERR_SHAPE;
  md2::ToTensorJobs jobs;
  unsigned blocks = 0;
  int used = 0;
  for (int i = 0; i < count; ++i) {
    const md2_u8_images& g = groups[i];
    if (g.N < 0 || g.H < 0 || g.W < 0) return MD2_ERR_SHAPE;
    const long long px = (long long)g.N * g.H * g.W;
    if (px == 0) continue;  // an empty group converts nothing
    if (!g.src || !g.dst) return MD2_ERR_NULL;
    if ((long long)g.H * g.W > 0x7fffffffLL || px > (1LL << 40)) return MD2_ERR_SHAPE;
    jobs.src[used] = g.src;
    jobs.dst[used] = g.dst;
    jobs.hw[used] = g.H * g.W;
    jobs.n[used] = g.N;
    jobs.first_block[used] = blocks;
    blocks += (unsigned)((px + 1023) / 1024);
    ++used;
  }
  if (used == 0) return 0;
  jobs.first_block[used] = blocks;
  jobs.count = used;
  md2::to_tensor_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(jobs);
  return (int)cudaGetLastError();
}

}  // extern "C"
