// md2_pipeline.cu - on-device colour pyramid: flip + Pillow-exact antialiased resize of the
// decoded uint8 image to every pyramid level + ToTensor, replacing kitti_mono.py:283-288,296-304,347-362.
//
// Host part: the Lanczos coefficient tables of Pillow's resampler (double precision, libm), one per axis
// and level.  Device part: two integer passes per level (horizontal to uint8, vertical to float / 255).
// Byte / integer work end to end; bound by HBM and the L1/L2 re-reads of the taps, no tensor cores.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/md2_pipeline.h"
#include "md2_host.h"

namespace md2 {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

// ---- table layout: per level, X axis then Y axis; per axis [bounds: 2 * out ints][coeffs: out * ksize ints]
struct AxisTab {
  int in, out, ksize;
  size_t offset;  // in ints from the start of the tables
};

static inline int lanczos_ksize(int in, int out) {
  double fs = (double)in / (double)out;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(3.0 * fs) * 2 + 1;
}

static inline size_t pyramid_layout(const md2_pyramid_cfg& c, AxisTab tab[kMaxScales][2]) {
  size_t off = 0;
  for (int s = 0; s < c.scales; ++s)
    for (int a = 0; a < 2; ++a) {
      AxisTab t;
      t.in = a == 0 ? c.Win : c.Hin;
      t.out = (a == 0 ? c.W : c.H) >> s;
      t.ksize = lanczos_ksize(t.in, t.out);
      t.offset = off;
      off += (size_t)t.out * (2 + t.ksize);
      if (tab) tab[s][a] = t;
    }
  return off;
}

static inline int validate_pyramid(const md2_pyramid_cfg* c) {
  if (!c) return MD2_ERR_NULL;
  if (c->N < 1 || c->Hin < 1 || c->Win < 1 || c->H < 1 || c->W < 1) return MD2_ERR_SHAPE;
  if (c->scales < 1 || c->scales > kMaxScales) return MD2_ERR_SHAPE;
  if ((c->H >> (c->scales - 1)) < 1 || (c->W >> (c->scales - 1)) < 1) return MD2_ERR_SHAPE;
  if ((long long)c->N * c->Hin * c->Win * 3 > 0x7fffffffLL) return MD2_ERR_SHAPE;
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

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis (box = the whole image)
static void fill_axis(const AxisTab& t, int* base) {
  int* bounds = base + t.offset;
  int* kk = bounds + 2 * (size_t)t.out;
  const double scale = (double)t.in / (double)t.out;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * fs, ss = 1.0 / fs;
  double w[1024];
  for (int xx = 0; xx < t.out; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > t.in) xmax = t.in;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = lanczos3((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = kk + (size_t)xx * t.ksize;
    for (int x = 0; x < t.ksize; ++x) {
      double v = x < xmax ? w[x] : 0.0;
      if (x < xmax && ww != 0.0) v /= ww;
      k[x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: [N, Hin, Win, 3] u8 -> [N, Hin, Wout, 3] u8
__global__ void __launch_bounds__(256) pyramid_h(int N, int Hin, int Win, int Wout, int ksize, const uint8_t* __restrict__ img,
                                                 const uint8_t* __restrict__ flip, const int* __restrict__ tab, uint8_t* tmp) {
  const int* bounds = tab;
  const int* kk = tab + 2 * Wout;
  const long long total = (long long)N * Hin * Wout;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int x = (int)(i % Wout);
    const long long row = i / Wout;  // n * Hin + y
    const int n = (int)(row / Hin);
    const int xmin = bounds[2 * x], cnt = bounds[2 * x + 1];
    const int* k = kk + (size_t)x * ksize;
    const uint8_t* src = img + row * (long long)Win * 3;
    const bool fl = flip && flip[n];
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < cnt; ++j) {
      const int xs = fl ? Win - 1 - (xmin + j) : xmin + j;
      const uint8_t* px = src + xs * 3;
      const int c = k[j];
      s0 += px[0] * c;
      s1 += px[1] * c;
      s2 += px[2] * c;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)clip8(s0);
    o[1] = (uint8_t)clip8(s1);
    o[2] = (uint8_t)clip8(s2);
  }
}

// vertical pass + ToTensor: [N, Hin, Wout, 3] u8 -> [N, 3, Hout, Wout] f32 (/255)
__global__ void __launch_bounds__(256) pyramid_v(int N, int Hin, int Hout, int Wout, int ksize, const uint8_t* __restrict__ tmp,
                                                 const int* __restrict__ tab, float* out) {
  const int* bounds = tab;
  const int* kk = tab + 2 * Hout;
  const long long total = (long long)N * Hout * Wout;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int x = (int)(i % Wout), y = (int)((i / Wout) % Hout), n = (int)(i / ((long long)Wout * Hout));
    const int ymin = bounds[2 * y], cnt = bounds[2 * y + 1];
    const int* k = kk + (size_t)y * ksize;
    const uint8_t* src = tmp + (((long long)n * Hin + ymin) * Wout + x) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < cnt; ++j) {
      const int c = k[j];
      s0 += src[0] * c;
      s1 += src[1] * c;
      s2 += src[2] * c;
      src += (long long)Wout * 3;
    }
    const long long plane = (long long)Hout * Wout;
    float* o = out + (long long)n * 3 * plane + (long long)y * Wout + x;
    o[0] = __fdiv_rn((float)clip8(s0), 255.0f);  // transforms.ToTensor: uint8 -> float32, .div(255)
    o[plane] = __fdiv_rn((float)clip8(s1), 255.0f);
    o[2 * plane] = __fdiv_rn((float)clip8(s2), 255.0f);
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
  AxisTab tab[kMaxScales][2];
  pyramid_layout(*cfg, tab);
  for (int s = 0; s < cfg->scales; ++s)
    for (int a = 0; a < 2; ++a) {
      if (tab[s][a].ksize > 1024) return MD2_ERR_SHAPE;
      fill_axis(tab[s][a], (int*)host_tables);
    }
  return 0;
}

size_t md2_pyramid_workspace_bytes(const md2_pyramid_cfg* cfg) {
  if (validate_pyramid(cfg) != 0) return 0;
  return (size_t)cfg->N * cfg->Hin * cfg->W * 3;  // level 0 is the widest intermediate
}

int md2_color_pyramid(const md2_pyramid_cfg* cfg, const uint8_t* images, const uint8_t* flip, const void* device_tables,
                      float* const* out, void* workspace, md2_stream_t stream) {
  const int v = validate_pyramid(cfg);
  if (v != 0) return v;
  if (!images || !device_tables || !out) return MD2_ERR_NULL;
  if (!workspace) return MD2_ERR_WORKSPACE;
  for (int s = 0; s < cfg->scales; ++s)
    if (!out[s]) return MD2_ERR_NULL;
  AxisTab tab[kMaxScales][2];
  pyramid_layout(*cfg, tab);
  cudaStream_t st = (cudaStream_t)stream;
  const int* tables = (const int*)device_tables;
  for (int s = 0; s < cfg->scales; ++s) {
    const AxisTab &tx = tab[s][0], &ty = tab[s][1];
    const long long nh = (long long)cfg->N * cfg->Hin * tx.out, nv = (long long)cfg->N * ty.out * tx.out;
    const int bh = (int)((nh + 255) / 256 > 148 * 16 ? 148 * 16 : (nh + 255) / 256);
    const int bv = (int)((nv + 255) / 256 > 148 * 16 ? 148 * 16 : (nv + 255) / 256);
    pyramid_h<<<bh, 256, 0, st>>>(cfg->N, cfg->Hin, cfg->Win, tx.out, tx.ksize, images, flip, tables + tx.offset,
                                  (uint8_t*)workspace);
    pyramid_v<<<bv, 256, 0, st>>>(cfg->N, cfg->Hin, ty.out, tx.out, ty.ksize, (const uint8_t*)workspace, tables + ty.offset,
                                  out[s]);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
