// md2_pool.cu - nn.MaxPool2d(k, stride, padding) of the ResNet encoders (model_layer/depth_encoder.py:29: 3, 2, 1),
// forward and backward, for channels-last tensors ("channels-last for the cuDNN nets").
//
// ATen's NHWC max-pool kernels are the slowest memory-bound operators left in the training step after the reflection
// pads (forward 0.34 ms + backward 0.59 ms of 15.2 ms at batch 12, 192x640: three encoder passes over
// [12|24, 64, 96, 320]; profiles/r2n_train_step_kernels.txt) - the backward scatters through atomics into a zero-filled
// buffer and both carry 8-byte indices.  Here:
//   forward   one thread per output pixel x 4 channels (float4): scans the window in ATen's order (rows, then columns;
//             a later element replaces the maximum only if it is greater or NaN, so ties go to the first), writes the
//             maximum and ONE BYTE per value, the winner's offset inside the window;
//   backward  a gather: every input element looks at the <= ceil(k/s)^2 windows that contain it and adds the output
//             gradients of those whose winner it is - no atomics, no zero fill, fixed order.
// Values are bit-identical to nn.MaxPool2d; gradients are equal up to the order of at most four additions.
// NHWC only (float4 when C % 4 == 0); the Python wrapper converts a tensor that is not channels-last first.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/md2_ops.h"
#include "md2_nvtx.h"

namespace md2 {

struct PoolShape {
  int C, H, W, Ho, Wo, k, s, p;
};

template <int VEC>
__global__ void __launch_bounds__(256) maxpool_fwd(const float* __restrict__ in, float* __restrict__ out,
                                                    uint8_t* __restrict__ win, const __grid_constant__ PoolShape q) {
  const int cv = q.C / VEC;
  const unsigned j = blockIdx.y * 256u + threadIdx.x;
  if (j >= (unsigned)(q.Wo * cv)) return;
  const int xo = j / cv, c = (j - xo * cv) * VEC;
  const unsigned row = blockIdx.x;  // n * Ho + yo
  const unsigned n = row / q.Ho;
  const int yo = row - n * q.Ho;
  const int y0 = yo * q.s - q.p, x0 = xo * q.s - q.p;
  const int ys = y0 < 0 ? 0 : y0, xs = x0 < 0 ? 0 : x0;
  const int ye = min(y0 + q.k, q.H), xe = min(x0 + q.k, q.W);
  float best[VEC];
  int where[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    best[v] = -INFINITY;
    where[v] = (ys - y0) * q.k + (xs - x0);  // ATen starts from the first in-bounds element
  }
  const float* base = in + (size_t)n * q.H * q.W * q.C + c;
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) {
      const float* ptr = base + ((size_t)y * q.W + x) * q.C;
      float val[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(ptr));
        val[0] = t.x; val[1 % VEC] = t.y; val[2 % VEC] = t.z; val[3 % VEC] = t.w;
      } else {
        val[0] = __ldg(ptr);
      }
      const int off = (y - y0) * q.k + (x - x0);
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (val[v] > best[v] || val[v] != val[v]) {
          best[v] = val[v];
          where[v] = off;
        }
    }
  const size_t o = ((size_t)row * q.Wo + xo) * q.C + c;
  if (VEC == 4) {
    *reinterpret_cast<float4*>(out + o) = make_float4(best[0], best[1 % VEC], best[2 % VEC], best[3 % VEC]);
    *reinterpret_cast<uchar4*>(win + o) = make_uchar4((unsigned char)where[0], (unsigned char)where[1 % VEC],
                                                       (unsigned char)where[2 % VEC], (unsigned char)where[3 % VEC]);
  } else {
    out[o] = best[0];
    win[o] = (uint8_t)where[0];
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) maxpool_bwd(const float* __restrict__ g_out, const uint8_t* __restrict__ win,
                                                    float* __restrict__ g_in, const __grid_constant__ PoolShape q) {
  const int cv = q.C / VEC;
  const unsigned j = blockIdx.y * 256u + threadIdx.x;
  if (j >= (unsigned)(q.W * cv)) return;
  const int x = j / cv, c = (j - x * cv) * VEC;
  const unsigned row = blockIdx.x;  // n * H + y
  const unsigned n = row / q.H;
  const int y = row - n * q.H;
  // output windows yo with yo*s - p <= y < yo*s - p + k
  const int yo_hi = min((y + q.p) / q.s, q.Ho - 1), xo_hi = min((x + q.p) / q.s, q.Wo - 1);
  const int ty = y + q.p - q.k + 1, tx = x + q.p - q.k + 1;
  const int yo_lo = ty <= 0 ? 0 : (ty + q.s - 1) / q.s, xo_lo = tx <= 0 ? 0 : (tx + q.s - 1) / q.s;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  for (int yo = yo_lo; yo <= yo_hi; ++yo)
    for (int xo = xo_lo; xo <= xo_hi; ++xo) {
      const int me = (y - (yo * q.s - q.p)) * q.k + (x - (xo * q.s - q.p));  // this element's offset in that window
      const size_t o = (((size_t)n * q.Ho + yo) * q.Wo + xo) * q.C + c;
      if (VEC == 4) {
        const uchar4 w = __ldg(reinterpret_cast<const uchar4*>(win + o));
        if (w.x == me || w.y == me || w.z == me || w.w == me) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(g_out + o));
          if (w.x == me) acc[0] += g.x;
          if (w.y == me) acc[1 % VEC] += g.y;
          if (w.z == me) acc[2 % VEC] += g.z;
          if (w.w == me) acc[3 % VEC] += g.w;
        }
      } else if (__ldg(win + o) == me) {
        acc[0] += __ldg(g_out + o);
      }
    }
  const size_t i = ((size_t)row * q.W + x) * q.C + c;
  if (VEC == 4) *reinterpret_cast<float4*>(g_in + i) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
  else g_in[i] = acc[0];
}

// The encoders' case (3, 2, 1) with float4 channels: one thread per 2 x 2 input block.  Row 2i lies in window row i only
// (offset 1), row 2i+1 in window rows i (offset 2) and i+1 (offset 0), the same along x: the four windows (i..i+1) x
// (j..j+1) are read once and serve four elements, each summed in the same window order as maxpool_bwd.
__global__ void __launch_bounds__(256) maxpool_bwd_3x3s2(const float* __restrict__ g_out, const uint8_t* __restrict__ win,
                                                          float* __restrict__ g_in, const __grid_constant__ PoolShape q) {
  const int cv = q.C / 4, Wb = (q.W + 1) / 2, Hb = (q.H + 1) / 2;
  const unsigned t = blockIdx.y * 256u + threadIdx.x;
  if (t >= (unsigned)(Wb * cv)) return;
  const int j = t / cv, c = (t - j * cv) * 4;
  const unsigned n = blockIdx.x / Hb;
  const int i = blockIdx.x - n * Hb;
  float4 g[2][2];
  uchar4 w[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int yo = i + a, xo = j + b;
      if (yo < q.Ho && xo < q.Wo) {
        const size_t o = (((size_t)n * q.Ho + yo) * q.Wo + xo) * q.C + c;
        w[a][b] = __ldg(reinterpret_cast<const uchar4*>(win + o));
        g[a][b] = __ldg(reinterpret_cast<const float4*>(g_out + o));
      } else {
        w[a][b] = make_uchar4(255, 255, 255, 255);  // no such window
        g[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  // element (2i + r, 2j + u) sits at offset (oy(r, a), ox(u, b)) of window (i + a, j + b): r = 0 -> a = 0 only, dy = 1;
  // r = 1 -> a = 0 with dy = 2, a = 1 with dy = 0
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int y = 2 * i + r, x = 2 * j + u;
      if (y >= q.H || x >= q.W) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a <= r; ++a)
#pragma unroll
        for (int b = 0; b <= u; ++b) {
          const int dy = r == 0 ? 1 : (a == 0 ? 2 : 0), dx = u == 0 ? 1 : (b == 0 ? 2 : 0);
          const unsigned char me = (unsigned char)(dy * 3 + dx);
          if (w[a][b].x == me) acc.x += g[a][b].x;
          if (w[a][b].y == me) acc.y += g[a][b].y;
          if (w[a][b].z == me) acc.z += g[a][b].z;
          if (w[a][b].w == me) acc.w += g[a][b].w;
        }
      *reinterpret_cast<float4*>(g_in + (((size_t)n * q.H + y) * q.W + x) * q.C + c) = acc;
    }
}

static int check_pool(int N, int C, int H, int W, int k, int s, int p, PoolShape* q) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || k <= 0 || s <= 0 || p < 0) return MD2_ERR_SHAPE;
  if (k > 15 || 2 * p > k) return MD2_ERR_SHAPE;  // offsets fit a byte; torch: "pad should be at most half of kernel size"
  const int Ho = (H + 2 * p - k) / s + 1, Wo = (W + 2 * p - k) / s + 1;  // ceil_mode = False, dilation 1
  if (H + 2 * p < k || W + 2 * p < k || Ho <= 0 || Wo <= 0) return MD2_ERR_SHAPE;
  if ((long long)N * H > 0x7fffffffLL || (long long)W * C > 0x7fffffffLL) return MD2_ERR_SHAPE;
  q->C = C; q->H = H; q->W = W; q->Ho = Ho; q->Wo = Wo; q->k = k; q->s = s; q->p = p;
  return 0;
}

}  // namespace md2

using namespace md2;

extern "C" {

int md2_maxpool2d_nhwc_forward(int N, int C, int H, int W, int kernel, int stride, int padding, const float* in,
                               float* out, uint8_t* winner, md2_stream_t stream) {
  PoolShape q;
  const int v = check_pool(N, C, H, W, kernel, stride, padding, &q);
  if (v != 0) return v;
  if (!in || !out || !winner) return MD2_ERR_NULL;
  const NvtxRange range("md2_maxpool2d_nhwc_forward");
  const bool vec = (C % 4 == 0) && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0) && (((uintptr_t)winner & 3) == 0);
  const int row_vectors = q.Wo * (vec ? C / 4 : C);
  const dim3 grid((unsigned)(N * q.Ho), (unsigned)((row_vectors + 255) / 256));
  if (grid.y > 65535) return MD2_ERR_SHAPE;
  if (vec) maxpool_fwd<4><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, winner, q);
  else maxpool_fwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, winner, q);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int md2_maxpool2d_nhwc_backward(int N, int C, int H, int W, int kernel, int stride, int padding, const float* g_out,
                                const uint8_t* winner, float* g_in, md2_stream_t stream) {
  PoolShape q;
  const int v = check_pool(N, C, H, W, kernel, stride, padding, &q);
  if (v != 0) return v;
  if (!g_out || !winner || !g_in) return MD2_ERR_NULL;
  const NvtxRange range("md2_maxpool2d_nhwc_backward");
  const bool vec = (C % 4 == 0) && ((((uintptr_t)g_out | (uintptr_t)g_in) & 15) == 0) && (((uintptr_t)winner & 3) == 0);
  if (vec && kernel == 3 && stride == 2 && padding == 1) {  // the encoders' stem
    const dim3 grid2((unsigned)(N * ((H + 1) / 2)), (unsigned)((((W + 1) / 2) * (C / 4) + 255) / 256));
    if (grid2.y > 65535) return MD2_ERR_SHAPE;
    maxpool_bwd_3x3s2<<<grid2, 256, 0, (cudaStream_t)stream>>>(g_out, winner, g_in, q);
    const cudaError_t e2 = cudaGetLastError();
    return e2 == cudaSuccess ? 0 : (int)e2;
  }
  const int row_vectors = q.W * (vec ? C / 4 : C);
  const dim3 grid((unsigned)(N * q.H), (unsigned)((row_vectors + 255) / 256));
  if (grid.y > 65535) return MD2_ERR_SHAPE;
  if (vec) maxpool_bwd<4><<<grid, 256, 0, (cudaStream_t)stream>>>(g_out, winner, g_in, q);
  else maxpool_bwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g_out, winner, g_in, q);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
