// md2_metrics.cu - training-time depth metrics (model_loss/model_metric.py:52-106) on sm_100a.
//
// Nine launches, no host sync, no allocation:
//   clear            zero the histograms and the select state
//   gather_hist      up-sample + clamp the prediction at every crop pixel, keep (gt, pred) of the masked pixels
//                    in the workspace, histogram the top 11 bits of both bit patterns
//   select x3        one block: find the histogram bin that holds the lower median, refine the key prefix
//   refine_hist x2   histogram the next 11 / 10 bits of the keys that match the prefix
//   accumulate       median scaling, second clamp, the seven error sums (fp64 accumulators per block)
//   finish           fixed-order reduction of the block partials -> out[8]
// The medians are exact (torch.median returns the element of rank (n-1)/2): positive fp32 values order like
// their bit patterns, so a radix select over 11 + 11 + 10 bits finds them without sorting.
#include <cuda_runtime.h>

#include "../../include/md2_metrics.h"
#include "md2_host.h"
#include "md2_nvtx.h"

namespace md2 {

constexpr int kBins = 2048;
constexpr int kAccBlocks = 296;  // 2 per SM

struct SelectState {
  unsigned prefix[2];  // key bits found so far (gt, pred)
  unsigned rank[2];    // rank of the median among the keys that share the prefix
  unsigned n;          // masked pixels
};

struct MetricsWs {
  float* gtv;
  float* prv;
  unsigned* hist;  // [3 passes][2][kBins]
  SelectState* st;
  double* part;    // [kAccBlocks][7]
};

MD2_HD long long crop_count(const md2_metrics_cfg& c) { return (long long)c.B * (c.y1 - c.y0) * (c.x1 - c.x0); }

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static inline size_t metrics_layout(const md2_metrics_cfg& c, char* base, MetricsWs* w) {
  const size_t n = (size_t)crop_count(c);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  char* a = take(n * 4);
  char* b = take(n * 4);
  char* h = take(3 * 2 * kBins * 4);
  char* s = take(sizeof(SelectState));
  char* p = take(kAccBlocks * 7 * 8);
  if (w) { w->gtv = (float*)a; w->prv = (float*)b; w->hist = (unsigned*)h; w->st = (SelectState*)s; w->part = (double*)p; }
  return off;
}

static inline int validate_metrics(const md2_metrics_cfg* c) {
  if (!c) return MD2_ERR_NULL;
  if (c->B < 1 || c->H < 1 || c->W < 1 || c->Hg < 1 || c->Wg < 1) return MD2_ERR_SHAPE;
  if (c->y0 < 0 || c->y1 > c->Hg || c->y0 >= c->y1 || c->x0 < 0 || c->x1 > c->Wg || c->x0 >= c->x1) return MD2_ERR_SHAPE;
  if (crop_count(*c) > 0x7fffffffLL) return MD2_ERR_SHAPE;
  if (!(c->min_depth > 0.f) || !(c->max_depth > c->min_depth)) return MD2_ERR_CONFIG;
  return 0;
}

__global__ void metrics_clear(unsigned* hist, SelectState* st) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * 2 * kBins; i += gridDim.x * blockDim.x) hist[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = SelectState{{0, 0}, {0, 0}, 0};
}

// ATen upsample_bilinear2d (align_corners = False, no scale factor): source index fma(scale, dst + 0.5, -0.5)
struct Src {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Src src_index(int v, float scale, int n_in) {
  float f = ffma(scale, (float)v + 0.5f, -0.5f);
  f = f < 0.f ? 0.f : f;
  Src o;
  o.i0 = imin((int)f, n_in - 1);
  o.i1 = o.i0 + (o.i0 < n_in - 1 ? 1 : 0);
  o.l1 = f - (float)o.i0;
  o.l0 = 1.0f - o.l1;
  return o;
}

__global__ void __launch_bounds__(256) metrics_gather_hist(md2_metrics_cfg c, float sy, float sx, const float* __restrict__ depth,
                                                           const float* __restrict__ gt, float* gtv, float* prv,
                                                           unsigned* hist) {
  __shared__ unsigned sh[2][kBins];
  for (int i = threadIdx.x; i < 2 * kBins; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int ch = c.y1 - c.y0, cw = c.x1 - c.x0;
  const long long n = crop_count(c);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const int xx = (int)(i % cw), yy = (int)((i / cw) % ch), b = (int)(i / ((long long)cw * ch));
    const int y = c.y0 + yy, x = c.x0 + xx;
    const float g = gt[((long long)b * c.Hg + y) * c.Wg + x];
    const bool valid = g > 0.f;
    float p = 0.f;
    if (valid) {
      const float* d = depth + (long long)b * c.H * c.W;
      const Src uy = src_index(y, sy, c.H), ux = src_index(x, sx, c.W);
      const float top = ffma(ux.l0, d[uy.i0 * c.W + ux.i0], fmul(ux.l1, d[uy.i0 * c.W + ux.i1]));
      const float bot = ffma(ux.l0, d[uy.i1 * c.W + ux.i0], fmul(ux.l1, d[uy.i1 * c.W + ux.i1]));
      p = ffma(uy.l0, top, fmul(uy.l1, bot));
      p = fminf(fmaxf(p, c.min_depth), c.max_depth);
      atomicAdd(&sh[0][__float_as_uint(g) >> 21], 1u);
      atomicAdd(&sh[1][__float_as_uint(p) >> 21], 1u);
    }
    gtv[i] = valid ? g : 0.f;
    prv[i] = p;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kBins; i += 256) {
    const unsigned v = (&sh[0][0])[i];
    if (v) atomicAdd(hist + i, v);
  }
}

// the keys whose upper bits equal the prefix found so far: histogram of their next `bits` bits
__global__ void __launch_bounds__(256) metrics_refine_hist(long long n, const float* __restrict__ gtv, const float* __restrict__ prv,
                                                           const SelectState* __restrict__ st, int shift, int bits,
                                                           unsigned* hist) {
  __shared__ unsigned sh[2][kBins];
  for (int i = threadIdx.x; i < 2 * kBins; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  const unsigned pg = st->prefix[0], pp = st->prefix[1], mask = (1u << bits) - 1u;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const float g = gtv[i];
    if (!(g > 0.f)) continue;
    const unsigned kg = __float_as_uint(g), kp = __float_as_uint(prv[i]);
    if ((kg >> (shift + bits)) == pg) atomicAdd(&sh[0][(kg >> shift) & mask], 1u);
    if ((kp >> (shift + bits)) == pp) atomicAdd(&sh[1][(kp >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kBins; i += 256) {
    const unsigned v = (&sh[0][0])[i];
    if (v) atomicAdd(hist + i, v);
  }
}

// one block of 1024 threads: locate the bin of the wanted rank in both histograms
__global__ void __launch_bounds__(1024) metrics_select(const unsigned* __restrict__ hist, SelectState* st, int pass, int bits) {
  __shared__ unsigned scan[1024];
  __shared__ unsigned total;
  const int t = threadIdx.x;
  for (int a = 0; a < 2; ++a) {
    const unsigned* h = hist + a * kBins;
    const unsigned h0 = h[2 * t], h1 = h[2 * t + 1];
    // read the state of the previous pass before the barriers of the scan: the winning thread overwrites it below
    const unsigned rank_in = pass == 0 ? 0u : st->rank[a];
    const unsigned prev = pass == 0 ? 0u : st->prefix[a];
    scan[t] = h0 + h1;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned v = t >= o ? scan[t - o] : 0u;
      __syncthreads();
      scan[t] += v;
      __syncthreads();
    }
    if (t == 1023) total = scan[1023];
    __syncthreads();
    unsigned rank;
    if (pass == 0) {
      if (t == 0 && a == 0) st->n = total;
      rank = total ? (total - 1u) / 2u : 0u;  // torch.median: lower median
    } else {
      rank = rank_in;
    }
    const unsigned before = scan[t] - (h0 + h1);
    if (total && rank >= before && rank < scan[t]) {
      const bool second = rank >= before + h0;
      const unsigned bin = 2u * t + (second ? 1u : 0u);
      st->prefix[a] = (prev << bits) | bin;
      st->rank[a] = rank - before - (second ? h0 : 0u);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) metrics_accumulate(long long n, const float* __restrict__ gtv, const float* __restrict__ prv,
                                                          const SelectState* __restrict__ st, float lo, float hi, double* part) {
  __shared__ double red[8][7];
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const float ratio = __fdiv_rn(__uint_as_float(st->prefix[0]), __uint_as_float(st->prefix[1]));
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const float g = gtv[i];
    if (!(g > 0.f)) continue;
    const float p = fminf(fmaxf(fmul(prv[i], ratio), lo), hi);
    const float thr = fmaxf(__fdiv_rn(g, p), __fdiv_rn(p, g));
    const float d = fsub(g, p), d2 = fmul(d, d);
    const float lg = fsub(logf(g), logf(p));
    acc[0] += (double)__fdiv_rn(fabsf(d), g);
    acc[1] += (double)__fdiv_rn(d2, g);
    acc[2] += (double)d2;
    acc[3] += (double)fmul(lg, lg);
    acc[4] += thr < 1.25f ? 1.0 : 0.0;
    acc[5] += thr < 1.5625f ? 1.0 : 0.0;
    acc[6] += thr < 1.953125f ? 1.0 : 0.0;
  }
#pragma unroll
  for (int e = 0; e < 7; ++e) {
    double v = acc[e];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][e] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    part[blockIdx.x * 7 + threadIdx.x] = v;
  }
}

__global__ void metrics_finish(int blocks, const double* __restrict__ part, const SelectState* __restrict__ st, float* out) {
  const int e = threadIdx.x;
  if (e < 7) {
    double v = 0;
    for (int b = 0; b < blocks; ++b) v += part[b * 7 + e];
    const double n = (double)st->n;
    double r = v / n;  // NaN when nothing is masked
    if (e == 2 || e == 3) r = sqrt(r);
    out[e] = (float)r;
  } else if (e == 7) {
    out[7] = (float)st->n;
  }
}

}  // namespace md2

using namespace md2;

extern "C" {

size_t md2_metrics_workspace_bytes(const md2_metrics_cfg* cfg) {
  if (validate_metrics(cfg) != 0) return 0;
  return metrics_layout(*cfg, nullptr, nullptr);
}

int md2_depth_metrics(const md2_metrics_cfg* cfg, const float* depth, const float* gt, float* out, void* workspace,
                      md2_stream_t stream) {
  const md2::NvtxRange range("md2_depth_metrics");
  const int v = validate_metrics(cfg);
  if (v != 0) return v;
  if (!depth || !gt || !out) return MD2_ERR_NULL;
  if (!workspace) return MD2_ERR_WORKSPACE;
  MetricsWs w;
  metrics_layout(*cfg, (char*)workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = crop_count(*cfg);
  int blocks = (int)((n + 255) / 256);
  blocks = blocks > 148 * 4 ? 148 * 4 : blocks;
  const float sy = (float)cfg->H / (float)cfg->Hg, sx = (float)cfg->W / (float)cfg->Wg;
  metrics_clear<<<8, 256, 0, st>>>(w.hist, w.st);
  metrics_gather_hist<<<blocks, 256, 0, st>>>(*cfg, sy, sx, depth, gt, w.gtv, w.prv, w.hist);
  metrics_select<<<1, 1024, 0, st>>>(w.hist, w.st, 0, 11);
  metrics_refine_hist<<<blocks, 256, 0, st>>>(n, w.gtv, w.prv, w.st, 10, 11, w.hist + 2 * kBins);
  metrics_select<<<1, 1024, 0, st>>>(w.hist + 2 * kBins, w.st, 1, 11);
  metrics_refine_hist<<<blocks, 256, 0, st>>>(n, w.gtv, w.prv, w.st, 0, 10, w.hist + 4 * kBins);
  metrics_select<<<1, 1024, 0, st>>>(w.hist + 4 * kBins, w.st, 2, 10);
  const int ab = blocks < kAccBlocks ? blocks : kAccBlocks;
  metrics_accumulate<<<ab, 256, 0, st>>>(n, w.gtv, w.prv, w.st, cfg->min_depth, cfg->max_depth, w.part);
  metrics_finish<<<1, 32, 0, st>>>(ab, w.part, w.st, out);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
