// md2_l1.cu - the reference's symbol-level (unfused) operators as stand-alone sm_100a kernels.
//
// SURVEY.md 8b "L1": interpolate, disparity2depth, Depth2PointCloud, PointCloud2Pixel, grid_sample
// (model_layer/warp.py) and ReprojectionLoss, SmoothLoss (model_loss/model_loss.py), each forward and
// backward, behind the same C ABI as the fused path.  They exist for drop-in completeness (a caller that
// composes the operators itself, e.g. the posecnn branch of processor.py:153-157); the training path uses
// the fused kernel of md2_abi.cu.  One thread per output element, the reference's rounding sequence
// (the same helpers as the fused tile code), no shared-memory tiling: these are compatibility symbols,
// not the hot path.
#include <cuda_runtime.h>

#include "../../include/md2_ops.h"
#include "md2_host.h"

namespace md2 {

#define MD2_GRID_STRIDE(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

static inline int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// --------------------------------------------------------------------------- disparity2depth (warp.py:29-39)
__global__ void disp2depth_fwd(long long n, const float* __restrict__ disp, float a, float r, float* scaled, float* depth) {
  MD2_GRID_STRIDE(i, n) {
    const float s = fadd(a, fmul(r, disp[i]));
    if (scaled) scaled[i] = s;
    if (depth) depth[i] = __frcp_rn(s);
  }
}
__global__ void disp2depth_bwd(long long n, const float* __restrict__ disp, float a, float r, const float* g_scaled,
                               const float* g_depth, float* g_disp) {
  MD2_GRID_STRIDE(i, n) {
    const float s = fadd(a, fmul(r, disp[i]));
    float g = g_scaled ? g_scaled[i] : 0.f;
    if (g_depth) g -= g_depth[i] / (s * s);  // d(1/s)/ds
    g_disp[i] = g * r;
  }
}

// --------------------------------------------------------------------------- interpolate (warp.py:18-20)
struct Up1 {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Up1 up1(int v, int n_in, int n_out) {
  // ATen area_pixel_compute_source_index, align_corners = False
  const float sc = (float)n_in / (float)n_out;
  float f = ffma(sc, (float)v + 0.5f, -0.5f);
  f = f < 0.f ? 0.f : f;
  Up1 o;
  o.i0 = imin((int)f, n_in - 1);
  o.i1 = o.i0 + (o.i0 < n_in - 1 ? 1 : 0);
  o.l1 = f - (float)o.i0;
  o.l0 = 1.0f - o.l1;
  return o;
}
__global__ void upsample_fwd(int planes, int h, int w, int H, int W, const float* __restrict__ in, float* out) {
  MD2_GRID_STRIDE(i, (long long)planes * H * W) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long pl = i / ((long long)H * W);
    const float* src = in + pl * h * w;
    if (h == H && w == W) {
      out[i] = src[y * w + x];
      continue;
    }
    const Up1 uy = up1(y, h, H), ux = up1(x, w, W);
    const float top = ffma(ux.l0, src[uy.i0 * w + ux.i0], fmul(ux.l1, src[uy.i0 * w + ux.i1]));
    const float bot = ffma(ux.l0, src[uy.i1 * w + ux.i0], fmul(ux.l1, src[uy.i1 * w + ux.i1]));
    out[i] = ffma(uy.l0, top, fmul(uy.l1, bot));
  }
}
__global__ void upsample_bwd(int planes, int h, int w, int H, int W, const float* __restrict__ g_out, float* g_in) {
  MD2_GRID_STRIDE(i, (long long)planes * H * W) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long pl = i / ((long long)H * W);
    float* dst = g_in + pl * h * w;
    const float g = g_out[i];
    const Up1 uy = up1(y, h, H), ux = up1(x, w, W);
    atomicAdd(dst + uy.i0 * w + ux.i0, g * uy.l0 * ux.l0);
    atomicAdd(dst + uy.i0 * w + ux.i1, g * uy.l0 * ux.l1);
    atomicAdd(dst + uy.i1 * w + ux.i0, g * uy.l1 * ux.l0);
    atomicAdd(dst + uy.i1 * w + ux.i1, g * uy.l1 * ux.l1);
  }
}
__global__ void fill_zero(long long n, float* p) {
  MD2_GRID_STRIDE(i, n) p[i] = 0.f;
}

// --------------------------------------------------------------------------- Depth2PointCloud (warp.py:193-246)
// mm: rounding of torch.matmul's tiny dot products (matmul_mode in md2_host.h)
__device__ __forceinline__ float mac_m(bool fma_mode, float a, float x, float acc) {
  return fma_mode ? ffma(a, x, acc) : fadd(acc, fmul(a, x));
}
__global__ void backproject_fwd(int B, int H, int W, const float* __restrict__ depth, const float* __restrict__ invK,
                                float* cam, int mm) {
  const long long N = (long long)H * W;
  MD2_GRID_STRIDE(i, (long long)B * N) {
    const int b = (int)(i / N);
    const long long pix = i - b * N;
    const float x = (float)(pix % W), y = (float)(pix / W);
    const float* k = invK + b * 16;
    const float d = depth[i];
    float* c = cam + (long long)b * 4 * N + pix;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = fmul(k[r * 4 + 0], x);
      acc = mac_m(mm == 0, k[r * 4 + 1], y, acc);
      acc = mac_m(mm == 0, k[r * 4 + 2], 1.0f, acc);
      c[r * N] = fmul(d, acc);
    }
    c[3 * N] = 1.0f;
  }
}
__global__ void backproject_bwd(int B, int H, int W, const float* __restrict__ invK, const float* __restrict__ g_cam,
                                float* g_depth) {
  const long long N = (long long)H * W;
  MD2_GRID_STRIDE(i, (long long)B * N) {
    const int b = (int)(i / N);
    const long long pix = i - b * N;
    const float x = (float)(pix % W), y = (float)(pix / W);
    const float* k = invK + b * 16;
    const float* g = g_cam + (long long)b * 4 * N + pix;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) acc += g[r * N] * (k[r * 4 + 0] * x + k[r * 4 + 1] * y + k[r * 4 + 2]);
    g_depth[i] = acc;
  }
}

// --------------------------------------------------------------------------- PointCloud2Pixel (warp.py:250-269)
__device__ __forceinline__ void load_P(const float* K, const float* T, bool kt_fma, float P[12]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = fmul(K[i * 4 + 0], T[0 * 4 + j]);
#pragma unroll
      for (int k = 1; k < 4; ++k) acc = mac_m(kt_fma, K[i * 4 + k], T[k * 4 + j], acc);
      P[i * 4 + j] = acc;
    }
}
__global__ void project_fwd(int B, int H, int W, const float* __restrict__ cam, const float* __restrict__ K,
                            const float* __restrict__ T, float eps, float inv_wm1, float inv_hm1, float* grid, int mm) {
  const long long N = (long long)H * W;
  MD2_GRID_STRIDE(i, (long long)B * N) {
    const int b = (int)(i / N);
    const long long pix = i - b * N;
    float P[12];
    load_P(K + b * 16, T + b * 16, B > 1, P);
    const float* c = cam + (long long)b * 4 * N + pix;
    const float c0 = c[0], c1 = c[N], c2 = c[2 * N], c3 = c[3 * N];
    float xyz[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = fmul(P[r * 4 + 0], c0);
      acc = mac_m(mm != 1, P[r * 4 + 1], c1, acc);
      acc = mac_m(mm != 1, P[r * 4 + 2], c2, acc);
      xyz[r] = mac_m(mm != 1, P[r * 4 + 3], c3, acc);
    }
    const float z = fadd(xyz[2], eps);
    const float u = __fdiv_rn(xyz[0], z), v = __fdiv_rn(xyz[1], z);
    grid[i * 2 + 0] = fmul(fsub(fmul(u, inv_wm1), 0.5f), 2.0f);
    grid[i * 2 + 1] = fmul(fsub(fmul(v, inv_hm1), 0.5f), 2.0f);
  }
}
__global__ void project_bwd(int B, int H, int W, const float* __restrict__ cam, const float* __restrict__ K,
                            const float* __restrict__ T, float eps, float inv_wm1, float inv_hm1,
                            const float* __restrict__ g_grid, float* g_cam, float* g_T) {
  // one block row per image: blockIdx.y = b, so that the dL/dP partial sums reduce per image
  const long long N = (long long)H * W;
  const int b = blockIdx.y;
  float P[12];
  load_P(K + b * 16, T + b * 16, true, P);
  float dP[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) dP[e] = 0.f;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < N; pix += (long long)gridDim.x * blockDim.x) {
    const float* c = cam + (long long)b * 4 * N + pix;
    const float cv[4] = {c[0], c[N], c[2 * N], c[3 * N]};
    float xyz[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) xyz[r] = P[r * 4] * cv[0] + P[r * 4 + 1] * cv[1] + P[r * 4 + 2] * cv[2] + P[r * 4 + 3] * cv[3];
    const float rz = 1.0f / (xyz[2] + eps);
    const float u = xyz[0] * rz, v = xyz[1] * rz;
    const float du = g_grid[((long long)b * N + pix) * 2 + 0] * 2.0f * inv_wm1;
    const float dv = g_grid[((long long)b * N + pix) * 2 + 1] * 2.0f * inv_hm1;
    const float d[3] = {du * rz, dv * rz, -(u * du + v * dv) * rz};
    float* gc = g_cam + (long long)b * 4 * N + pix;
#pragma unroll
    for (int j = 0; j < 4; ++j) gc[j * N] = d[0] * P[j] + d[1] * P[4 + j] + d[2] * P[8 + j];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) dP[r * 4 + j] += d[r] * cv[j];
  }
  // dL/dT = K^T [dP; 0]; block reduction then one atomic per entry
  __shared__ float red[8][12];
#pragma unroll
  for (int e = 0; e < 12; ++e) {
    const float x = warp_sum(dP[e]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][e] = x;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int k = threadIdx.x >> 2, j = threadIdx.x & 3;
    float acc = 0.f;
    for (int r = 0; r < 3; ++r) {
      float s = 0.f;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += red[wv][r * 4 + j];
      acc += K[b * 16 + r * 4 + k] * s;
    }
    atomicAdd(g_T + b * 16 + k * 4 + j, acc);
  }
}

// --------------------------------------------------------------------------- grid_sample (warp.py:12-14)
// bilinear, padding_mode="border", align_corners=True (ATen GridSampler.cuh)
struct GS {
  int x0, y0, x1, y1;
  float wnw, wne, wsw, wse, ax, ay, bx, by;
  bool mx, my;
};
__device__ __forceinline__ GS gs_setup(float gx, float gy, int H, int W) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float ix = fmul(fmul(fadd(gx, 1.0f), 0.5f), wm1), iy = fmul(fmul(fadd(gy, 1.0f), 0.5f), hm1);
  GS g;
  g.mx = (ix > 0.0f) && (ix < wm1);
  g.my = (iy > 0.0f) && (iy < hm1);
  ix = fminf(wm1, fmaxf(ix, 0.0f));
  iy = fminf(hm1, fmaxf(iy, 0.0f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  g.ax = fsub(ix, x0f); g.ay = fsub(iy, y0f);
  g.bx = fsub(fadd(x0f, 1.0f), ix); g.by = fsub(fadd(y0f, 1.0f), iy);
  g.x0 = (int)x0f; g.y0 = (int)y0f;
  g.x1 = imin(g.x0 + 1, W - 1); g.y1 = imin(g.y0 + 1, H - 1);  // the weight of a clamped corner is 0
  g.wnw = fmul(g.bx, g.by); g.wne = fmul(g.ax, g.by); g.wsw = fmul(g.bx, g.ay); g.wse = fmul(g.ax, g.ay);
  return g;
}
__global__ void grid_sample_fwd(int B, int C, int H, int W, int Ho, int Wo, const float* __restrict__ img,
                                const float* __restrict__ grid, float* out) {
  const long long No = (long long)Ho * Wo;
  MD2_GRID_STRIDE(i, (long long)B * No) {
    const int b = (int)(i / No);
    const long long pix = i - b * No;
    const GS g = gs_setup(grid[i * 2], grid[i * 2 + 1], H, W);
    for (int c = 0; c < C; ++c) {
      const float* pl = img + ((long long)b * C + c) * H * W;
      float acc = fmul(pl[g.y0 * W + g.x0], g.wnw);
      acc = ffma(pl[g.y0 * W + g.x1], g.wne, acc);
      acc = ffma(pl[g.y1 * W + g.x0], g.wsw, acc);
      acc = ffma(pl[g.y1 * W + g.x1], g.wse, acc);
      out[((long long)b * C + c) * No + pix] = acc;
    }
  }
}
__global__ void grid_sample_bwd(int B, int C, int H, int W, int Ho, int Wo, const float* __restrict__ img,
                                const float* __restrict__ grid, const float* __restrict__ g_out, float* g_grid) {
  const long long No = (long long)Ho * Wo;
  MD2_GRID_STRIDE(i, (long long)B * No) {
    const int b = (int)(i / No);
    const long long pix = i - b * No;
    const GS g = gs_setup(grid[i * 2], grid[i * 2 + 1], H, W);
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* pl = img + ((long long)b * C + c) * H * W;
      const float go = g_out[((long long)b * C + c) * No + pix];
      const float vnw = pl[g.y0 * W + g.x0], vne = pl[g.y0 * W + g.x1], vsw = pl[g.y1 * W + g.x0], vse = pl[g.y1 * W + g.x1];
      gix += go * ((vne - vnw) * g.by + (vse - vsw) * g.ay);
      giy += go * ((vsw - vnw) * g.bx + (vse - vne) * g.ax);
    }
    // d ix / d gx = (W-1)/2, zero where the coordinate was clipped
    g_grid[i * 2 + 0] = g.mx ? gix * 0.5f * (float)(W - 1) : 0.f;
    g_grid[i * 2 + 1] = g.my ? giy * 0.5f * (float)(H - 1) : 0.f;
  }
}

// --------------------------------------------------------------------------- ReprojectionLoss (model_loss.py:92-103)
__device__ __forceinline__ int refl(int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }

template <bool BWD>
__device__ __forceinline__ float reproj_window(const float* __restrict__ pred, const float* __restrict__ tgt, int H, int W,
                                               int y, int x, float c1, float c2, float cf[9]) {
  float ss = 0.f, l1 = 0.f;
  const long long HW = (long long)H * W;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float xv[9], xx[9], xy[9], yv[9], yy[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int k = (dy + 1) * 3 + dx + 1;
        const long long o = ch * HW + (long long)refl(y + dy, H) * W + refl(x + dx, W);
        const float a = pred[o], t = tgt[o];
        xv[k] = a; yv[k] = t;
        xx[k] = fmul(a, a); yy[k] = fmul(t, t); xy[k] = fmul(a, t);
      }
    const float mu_y = div9(sum9(yv)), e2_y = div9(sum9(yy));
    const float sv = ssim_from_sums<BWD>(sum9(xv), sum9(xx), sum9(xy), mu_y, e2_y, c1, c2, cf[ch * 3], cf[ch * 3 + 1], cf[ch * 3 + 2]);
    const float lv = fabsf(fsub(yv[4], xv[4]));
    ss = ch == 0 ? sv : fadd(ss, sv);
    l1 = ch == 0 ? lv : fadd(l1, lv);
  }
  return fadd(fmul(0.85f, fmul(ss, kThird)), fmul(0.15f, fmul(l1, kThird)));
}
__global__ void reprojection_fwd(int B, int H, int W, const float* __restrict__ pred, const float* __restrict__ tgt, float* out) {
  const long long HW = (long long)H * W;
  MD2_GRID_STRIDE(i, (long long)B * HW) {
    const int b = (int)(i / HW);
    const long long pix = i - b * HW;
    float cf[9];
    out[i] = reproj_window<false>(pred + (long long)b * 3 * HW, tgt + (long long)b * 3 * HW, H, W, (int)(pix / W), (int)(pix % W),
                                  (float)0.0001, (float)0.0009, cf);
  }
}
// gradient wrt the prediction: every pixel gathers from the (up to 9, reflection-weighted) windows that contain it
__global__ void reprojection_bwd(int B, int H, int W, const float* __restrict__ pred, const float* __restrict__ tgt,
                                 const float* __restrict__ g_out, float* g_pred) {
  const long long HW = (long long)H * W;
  MD2_GRID_STRIDE(i, (long long)B * HW) {
    const int b = (int)(i / HW);
    const long long pix = i - b * HW;
    const int y = (int)(pix / W), x = (int)(pix % W);
    const float* pb = pred + (long long)b * 3 * HW;
    const float* tb = tgt + (long long)b * 3 * HW;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int qy = y + dy, qx = x + dx;
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
        // multiplicity of pixel (y, x) among the reflected taps of window (qy, qx)
        const float wy = (dy == -1 && y == 1) || (dy == 1 && y == H - 2) ? 2.f : 1.f;
        const float wx = (dx == -1 && x == 1) || (dx == 1 && x == W - 2) ? 2.f : 1.f;
        float cf[9];
        (void)reproj_window<true>(pb, tb, H, W, qy, qx, (float)0.0001, (float)0.0009, cf);
        const float h = g_out[(long long)b * HW + (long long)qy * W + qx] * (0.85f / 3.0f) * (-0.5f) * wy * wx;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          acc[ch] += h * (cf[ch * 3] + cf[ch * 3 + 1] * tb[ch * HW + pix] + cf[ch * 3 + 2] * pb[ch * HW + pix]);
      }
    const float gl1 = g_out[i] * (0.15f / 3.0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float w = pb[ch * HW + pix], t = tb[ch * HW + pix];
      g_pred[((long long)b * 3 + ch) * HW + pix] = acc[ch] + (w > t ? gl1 : (w < t ? -gl1 : 0.f));
    }
  }
}

// --------------------------------------------------------------------------- SmoothLoss (model_loss.py:107-116)
// one block per image: sums (mean, |dx|e, |dy|e) -> part[b][3]; a second tiny kernel combines them
__global__ void __launch_bounds__(256) smooth_l1_sums(int h, int w, const float* __restrict__ disp, const float* __restrict__ color,
                                                      float* part) {
  __shared__ float red[8][3];
  const int b = blockIdx.x, n = h * w;
  const float* d = disp + (long long)b * n;
  const float* col = color + (long long)b * 3 * n;
  float v[3] = {0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += 256) {
    const int y = i / w, x = i - y * w;
    const float di = d[i];
    v[0] += di;
    if (x + 1 < w) v[1] += fabsf(di - d[i + 1]) * edge_weight(col, n, i, i + 1);
    if (y + 1 < h) v[2] += fabsf(di - d[i + w]) * edge_weight(col, n, i, i + w);
  }
  for (int e = 0; e < 3; ++e) {
    const float x = warp_sum(v[e]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][e] = x;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += red[wv][threadIdx.x];
    part[b * 3 + threadIdx.x] = a;
  }
}
__global__ void smooth_l1_loss(int B, int h, int w, const float* __restrict__ part, float* loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) {
      const double inv = 1.0 / ((double)part[b * 3] / ((double)h * w) + 1e-7);
      if (w > 1) acc += inv * part[b * 3 + 1] / ((double)B * h * (w - 1));
      if (h > 1) acc += inv * part[b * 3 + 2] / ((double)B * (h - 1) * w);
    }
    *loss = (float)acc;
  }
}
__global__ void __launch_bounds__(256) smooth_l1_bwd(int B, int h, int w, const float* __restrict__ disp,
                                                     const float* __restrict__ color, const float* __restrict__ part,
                                                     const float* __restrict__ g_loss, float* g_disp) {
  const int n = h * w;
  const float gl = g_loss[0];
  MD2_GRID_STRIDE(i, (long long)B * n) {
    const int b = (int)(i / n), j = (int)(i - (long long)b * n);
    const int y = j / w, x = j - y * w;
    const float* d = disp + (long long)b * n;
    const float* col = color + (long long)b * 3 * n;
    const float inv = 1.0f / (part[b * 3] / (float)n + 1e-7f);
    const float cx = w > 1 ? gl / ((float)B * h * (w - 1)) : 0.f, cy = h > 1 ? gl / ((float)B * (h - 1) * w) : 0.f;
    const float Lb = inv * (cx * part[b * 3 + 1] + cy * part[b * 3 + 2]);
    const float di = d[j];
    float dn = 0.f;
    if (x + 1 < w) dn += cx * sgn(di - d[j + 1]) * edge_weight(col, n, j, j + 1);
    if (x > 0) dn -= cx * sgn(d[j - 1] - di) * edge_weight(col, n, j - 1, j);
    if (y + 1 < h) dn += cy * sgn(di - d[j + w]) * edge_weight(col, n, j, j + w);
    if (y > 0) dn -= cy * sgn(d[j - w] - di) * edge_weight(col, n, j - w, j);
    g_disp[i] = dn * inv - Lb * inv / (float)n;
  }
}

// --------------------------------------------------------------------------- mean(1 / depth) per image (processor.py:155)
__global__ void __launch_bounds__(256) mean_inv_fwd(int n, const float* __restrict__ depth, float* out) {
  __shared__ float red[8];
  const float* d = depth + (long long)blockIdx.x * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += __frcp_rn(d[i]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += red[wv];
    out[blockIdx.x] = a / (float)n;
  }
}
__global__ void mean_inv_bwd(int B, int n, const float* __restrict__ depth, const float* __restrict__ g, float* g_depth) {
  MD2_GRID_STRIDE(i, (long long)B * n) {
    const float d = depth[i];
    g_depth[i] = -g[i / n] / ((float)n * d * d);
  }
}

static inline int rc(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

}  // namespace md2

using namespace md2;

extern "C" {

int md2_disp2depth_forward(long long n, const float* disp, double min_depth, double max_depth, float* scaled, float* depth,
                           md2_stream_t st) {
  if (n <= 0 || !(min_depth > 0.0) || !(max_depth > min_depth)) return MD2_ERR_SHAPE;
  if (!disp || (!scaled && !depth)) return MD2_ERR_NULL;
  const float a = (float)(1.0 / max_depth), r = (float)(1.0 / min_depth - 1.0 / max_depth);
  disp2depth_fwd<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, disp, a, r, scaled, depth);
  return rc(cudaGetLastError());
}
int md2_disp2depth_backward(long long n, const float* disp, double min_depth, double max_depth, const float* g_scaled,
                            const float* g_depth, float* g_disp, md2_stream_t st) {
  if (n <= 0 || !(min_depth > 0.0) || !(max_depth > min_depth)) return MD2_ERR_SHAPE;
  if (!disp || !g_disp) return MD2_ERR_NULL;
  const float a = (float)(1.0 / max_depth), r = (float)(1.0 / min_depth - 1.0 / max_depth);
  disp2depth_bwd<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, disp, a, r, g_scaled, g_depth, g_disp);
  return rc(cudaGetLastError());
}

int md2_upsample_forward(int planes, int h, int w, int H, int W, const float* in, float* out, md2_stream_t st) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MD2_ERR_SHAPE;
  if (!in || !out) return MD2_ERR_NULL;
  upsample_fwd<<<blocks_for((long long)planes * H * W), 256, 0, (cudaStream_t)st>>>(planes, h, w, H, W, in, out);
  return rc(cudaGetLastError());
}
int md2_upsample_backward(int planes, int h, int w, int H, int W, const float* g_out, float* g_in, md2_stream_t st) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MD2_ERR_SHAPE;
  if (!g_out || !g_in) return MD2_ERR_NULL;
  fill_zero<<<blocks_for((long long)planes * h * w), 256, 0, (cudaStream_t)st>>>((long long)planes * h * w, g_in);
  upsample_bwd<<<blocks_for((long long)planes * H * W), 256, 0, (cudaStream_t)st>>>(planes, h, w, H, W, g_out, g_in);
  return rc(cudaGetLastError());
}

int md2_backproject_forward(int B, int H, int W, const float* depth, const float* inv_K, float* cam, md2_stream_t st) {
  if (B <= 0 || H <= 0 || W <= 0) return MD2_ERR_SHAPE;
  if (!depth || !inv_K || !cam) return MD2_ERR_NULL;
  backproject_fwd<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)st>>>(B, H, W, depth, inv_K, cam, matmul_mode(B, H, W));
  return rc(cudaGetLastError());
}
int md2_backproject_backward(int B, int H, int W, const float* inv_K, const float* g_cam, float* g_depth, md2_stream_t st) {
  if (B <= 0 || H <= 0 || W <= 0) return MD2_ERR_SHAPE;
  if (!inv_K || !g_cam || !g_depth) return MD2_ERR_NULL;
  backproject_bwd<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)st>>>(B, H, W, inv_K, g_cam, g_depth);
  return rc(cudaGetLastError());
}

int md2_project_forward(int B, int H, int W, const float* cam, const float* K, const float* T, double eps, float* grid,
                        md2_stream_t st) {
  if (B <= 0 || H <= 1 || W <= 1) return MD2_ERR_SHAPE;
  if (!cam || !K || !T || !grid) return MD2_ERR_NULL;
  project_fwd<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)st>>>(B, H, W, cam, K, T, (float)eps, 1.0f / (float)(W - 1),
                                                                               1.0f / (float)(H - 1), grid, matmul_mode(B, H, W));
  return rc(cudaGetLastError());
}
int md2_project_backward(int B, int H, int W, const float* cam, const float* K, const float* T, double eps, const float* g_grid,
                         float* g_cam, float* g_T, md2_stream_t st) {
  if (B <= 0 || H <= 1 || W <= 1) return MD2_ERR_SHAPE;
  if (!cam || !K || !T || !g_grid || !g_cam || !g_T) return MD2_ERR_NULL;
  fill_zero<<<1, 256, 0, (cudaStream_t)st>>>((long long)B * 16, g_T);
  int bx = (int)(((long long)H * W + 255) / 256);
  bx = bx > 120 ? 120 : bx;
  project_bwd<<<dim3(bx, B), 256, 0, (cudaStream_t)st>>>(B, H, W, cam, K, T, (float)eps, 1.0f / (float)(W - 1), 1.0f / (float)(H - 1),
                                                         g_grid, g_cam, g_T);
  return rc(cudaGetLastError());
}

int md2_grid_sample_forward(int B, int C, int H, int W, int Ho, int Wo, const float* img, const float* grid, float* out,
                            md2_stream_t st) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return MD2_ERR_SHAPE;
  if (!img || !grid || !out) return MD2_ERR_NULL;
  grid_sample_fwd<<<blocks_for((long long)B * Ho * Wo), 256, 0, (cudaStream_t)st>>>(B, C, H, W, Ho, Wo, img, grid, out);
  return rc(cudaGetLastError());
}
int md2_grid_sample_backward(int B, int C, int H, int W, int Ho, int Wo, const float* img, const float* grid,
                             const float* g_out, float* g_grid, md2_stream_t st) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return MD2_ERR_SHAPE;
  if (!img || !grid || !g_out || !g_grid) return MD2_ERR_NULL;
  grid_sample_bwd<<<blocks_for((long long)B * Ho * Wo), 256, 0, (cudaStream_t)st>>>(B, C, H, W, Ho, Wo, img, grid, g_out, g_grid);
  return rc(cudaGetLastError());
}

int md2_reprojection_forward(int B, int H, int W, const float* pred, const float* target, float* out, md2_stream_t st) {
  if (B <= 0 || H < 3 || W < 3) return MD2_ERR_SHAPE;
  if (!pred || !target || !out) return MD2_ERR_NULL;
  reprojection_fwd<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)st>>>(B, H, W, pred, target, out);
  return rc(cudaGetLastError());
}
int md2_reprojection_backward(int B, int H, int W, const float* pred, const float* target, const float* g_out, float* g_pred,
                              md2_stream_t st) {
  if (B <= 0 || H < 3 || W < 3) return MD2_ERR_SHAPE;
  if (!pred || !target || !g_out || !g_pred) return MD2_ERR_NULL;
  reprojection_bwd<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)st>>>(B, H, W, pred, target, g_out, g_pred);
  return rc(cudaGetLastError());
}

int md2_smooth_forward(int B, int h, int w, const float* disp, const float* color, float* loss, float* part, md2_stream_t st) {
  if (B <= 0 || h <= 0 || w <= 0) return MD2_ERR_SHAPE;
  if (!disp || !color || !loss || !part) return MD2_ERR_NULL;
  smooth_l1_sums<<<B, 256, 0, (cudaStream_t)st>>>(h, w, disp, color, part);
  smooth_l1_loss<<<1, 32, 0, (cudaStream_t)st>>>(B, h, w, part, loss);
  return rc(cudaGetLastError());
}
int md2_smooth_backward(int B, int h, int w, const float* disp, const float* color, const float* part, const float* g_loss_dev,
                        float* g_disp, md2_stream_t st) {
  if (B <= 0 || h <= 0 || w <= 0) return MD2_ERR_SHAPE;
  if (!disp || !color || !part || !g_loss_dev || !g_disp) return MD2_ERR_NULL;
  smooth_l1_bwd<<<blocks_for((long long)B * h * w), 256, 0, (cudaStream_t)st>>>(B, h, w, disp, color, part, g_loss_dev, g_disp);
  return rc(cudaGetLastError());
}

int md2_mean_inv_depth_forward(int B, int n, const float* depth, float* out, md2_stream_t st) {
  if (B <= 0 || n <= 0) return MD2_ERR_SHAPE;
  if (!depth || !out) return MD2_ERR_NULL;
  mean_inv_fwd<<<B, 256, 0, (cudaStream_t)st>>>(n, depth, out);
  return rc(cudaGetLastError());
}
int md2_mean_inv_depth_backward(int B, int n, const float* depth, const float* g_out, float* g_depth, md2_stream_t st) {
  if (B <= 0 || n <= 0) return MD2_ERR_SHAPE;
  if (!depth || !g_out || !g_depth) return MD2_ERR_NULL;
  mean_inv_bwd<<<blocks_for((long long)B * n), 256, 0, (cudaStream_t)st>>>(B, n, depth, g_out, g_depth);
  return rc(cudaGetLastError());
}

}  // extern "C"
