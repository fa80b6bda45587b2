// md2_jitter.cu - colour jitter on the device (do_color branch of kitti_mono.py:351-357),
// bit-identical to torchvision's PIL path: Pillow's Image.blend (fp32 product + sum, truncation), rgb2l (16-bit fixed
// point luma), rgb2hsv_row / hsv2rgb (fp32 ratios, hue through double, C round()).  Byte work: every step reads
// and writes a planar uint8 copy [N,3,H,W]; one launch per adjustment (+ an exact integer luma sum for contrast).
#include <cuda_runtime.h>

#include "../../include/md2_pipeline.h"
#include "md2_host.h"
#include "md2_nvtx.h"

namespace md2 {

__device__ __forceinline__ float jit_div255(int v) {
  const float r = 1.0f / 255.0f, x = (float)v;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-255.0f, q, x), r, q);
}

// Image.blend(im1, im2, alpha) for one byte (libImaging/Blend.c)
__device__ __forceinline__ int jit_blend(int a, int b, float alpha, bool inside) {
  const float t = __fadd_rn((float)a, __fmul_rn(alpha, (float)(b - a)));
  if (inside) return (int)t;  // alpha in [0, 1]: plain truncation
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}

__device__ __forceinline__ int jit_luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Convert.c rgb2hsv_row
__device__ __forceinline__ void jit_rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) {
    uh = 0;
    us = 0;
    return;
  }
  const float cr = (float)(maxc - minc);
  const float s = __fdiv_rn(cr, (float)maxc);
  const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
  else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
  double x = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);  // in (0.5, 2): fmod(x, 1) = x - floor(x), exact
  h = (float)(x - floor(x));
  const int ih = (int)__dmul_rn((double)h, 255.0), is = (int)__dmul_rn((double)s, 255.0);
  uh = min(max(ih, 0), 255);
  us = min(max(is, 0), 255);
}

// Convert.c hsv2rgb
__device__ __forceinline__ void jit_hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) {
    r = g = b = v;
    return;
  }
  const double hf = __ddiv_rn(__dmul_rn((double)h, 6.0), 255.0);
  const int i = (int)floor(hf);
  const double f = (double)(float)__dsub_rn(hf, (double)i);
  const double fs = (double)(float)__ddiv_rn((double)s, 255.0);
  const double vf = (double)v;
  const int p = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, fs))), 0), 255);
  const int q = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f)))), 0), 255);
  const int t = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f))))), 0), 255);
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// grid (pixels / 256, N).  float planes (v/255) -> uint8 planes; exact: v/255 is correctly rounded, so *255 rounds back
__global__ void __launch_bounds__(256) jit_to_u8(int hw, const float* __restrict__ in, uint8_t* out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  const size_t base = (size_t)blockIdx.y * 3 * hw + i;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[base + (size_t)c * hw] = (uint8_t)__float2int_rn(__fmul_rn(in[base + (size_t)c * hw], 255.0f));
}

__global__ void __launch_bounds__(256) jit_to_float(int hw, const uint8_t* __restrict__ jit, const float* __restrict__ in,
                                                    const uint8_t* __restrict__ apply, float* out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  const size_t base = (size_t)blockIdx.y * 3 * hw + i;
  const bool on = !apply || apply[blockIdx.y];
#pragma unroll
  for (int c = 0; c < 3; ++c) out[base + (size_t)c * hw] = on ? jit_div255(jit[base + (size_t)c * hw]) : in[base + (size_t)c * hw];
}

// exact integer sum of the luma of every image (ImageStat.Stat(img.convert("L")).mean)
__global__ void __launch_bounds__(256) jit_luma_sum(int hw, const uint8_t* __restrict__ img, unsigned long long* sums) {
  __shared__ unsigned red[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  unsigned v = 0;
  if (i < hw) {
    const size_t base = (size_t)blockIdx.y * 3 * hw + i;
    v = (unsigned)jit_luma(img[base], img[base + hw], img[base + 2 * (size_t)hw]);
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(sums + blockIdx.y, (unsigned long long)t);
  }
}

// one adjustment, in place on the uint8 planes.  op: 0 brightness, 1 contrast, 2 saturation, 3 hue
__global__ void __launch_bounds__(256) jit_adjust(int hw, int op, float factor, int hue_shift, uint8_t* img,
                                                  const unsigned long long* __restrict__ sums) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  const size_t base = (size_t)blockIdx.y * 3 * hw + i;
  int r = img[base], g = img[base + hw], b = img[base + 2 * (size_t)hw];
  const bool inside = factor >= 0.f && factor <= 1.f;
  if (op == 0) {
    r = jit_blend(0, r, factor, inside); g = jit_blend(0, g, factor, inside); b = jit_blend(0, b, factor, inside);
  } else if (op == 1) {
    const int mean = (int)((double)sums[blockIdx.y] / (double)hw + 0.5);
    r = jit_blend(mean, r, factor, inside); g = jit_blend(mean, g, factor, inside); b = jit_blend(mean, b, factor, inside);
  } else if (op == 2) {
    const int L = jit_luma(r, g, b);
    r = jit_blend(L, r, factor, inside); g = jit_blend(L, g, factor, inside); b = jit_blend(L, b, factor, inside);
  } else {
    int h, s, v;
    jit_rgb2hsv(r, g, b, h, s, v);
    jit_hsv2rgb((h + hue_shift) & 255, s, v, r, g, b);
  }
  img[base] = (uint8_t)r;
  img[base + hw] = (uint8_t)g;
  img[base + 2 * (size_t)hw] = (uint8_t)b;
}

static inline int validate_jitter(const md2_jitter_cfg* c) {
  if (!c) return MD2_ERR_NULL;
  if (c->N < 1 || c->N > 65535 || c->H < 1 || c->W < 1 || (long long)c->H * c->W > 0x3fffffffLL) return MD2_ERR_SHAPE;
  for (int k = 0; k < 4; ++k)
    if (c->order[k] < -1 || c->order[k] > 3) return MD2_ERR_CONFIG;
  if (!(c->brightness >= 0.0) || !(c->contrast >= 0.0) || !(c->saturation >= 0.0) || !(c->hue >= -0.5 && c->hue <= 0.5))
    return MD2_ERR_CONFIG;
  return 0;
}

}  // namespace md2

using namespace md2;

extern "C" {

size_t md2_jitter_workspace_bytes(const md2_jitter_cfg* cfg) {
  if (validate_jitter(cfg) != 0) return 0;
  const size_t img = (((size_t)cfg->N * 3 * cfg->H * cfg->W + 255) / 256) * 256;
  return img + (((size_t)cfg->N * 8 + 255) / 256) * 256;
}

int md2_color_jitter(const md2_jitter_cfg* cfg, const float* in, const uint8_t* apply, float* out, void* workspace,
                     md2_stream_t stream) {
  const md2::NvtxRange range("md2_color_jitter");
  const int v = validate_jitter(cfg);
  if (v != 0) return v;
  if (!in || !out) return MD2_ERR_NULL;
  if (!workspace) return MD2_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = cfg->H * cfg->W;
  const size_t img_bytes = (((size_t)cfg->N * 3 * hw + 255) / 256) * 256;
  uint8_t* img = (uint8_t*)workspace;
  unsigned long long* sums = (unsigned long long*)((uint8_t*)workspace + img_bytes);
  const dim3 grid((hw + 255) / 256, cfg->N);
  jit_to_u8<<<grid, 256, 0, st>>>(hw, in, img);
  // Image.blend takes its factor as a C float
  const float factor[4] = {(float)cfg->brightness, (float)cfg->contrast, (float)cfg->saturation, (float)cfg->hue};
  // functional_pil.adjust_hue adds np.int32(hue * 255).astype(np.uint8) to the H channel (wrap-around)
  const int hue_shift = (int)(cfg->hue * 255.0) & 255;
  for (int k = 0; k < 4; ++k) {
    const int op = cfg->order[k];
    if (op < 0) continue;
    if (op == 1) {
      cudaMemsetAsync(sums, 0, (size_t)cfg->N * 8, st);
      jit_luma_sum<<<grid, 256, 0, st>>>(hw, img, sums);
    }
    jit_adjust<<<grid, 256, 0, st>>>(hw, op, factor[op], hue_shift, img, sums);
  }
  jit_to_float<<<grid, 256, 0, st>>>(hw, img, in, apply, out);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
