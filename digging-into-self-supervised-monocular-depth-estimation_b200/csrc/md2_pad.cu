// md2_pad.cu - nn.ReflectionPad2d for the decoder's Conv3x3 blocks (model_layer/depth_decoder.py:36-50,
// model_layer/warp.py:175-190), forward and backward, in the memory format the tensor already has.
//
// Why it exists ("channels-last for the cuDNN nets"): ATen's reflection_pad2d only knows NCHW.  In a
// channels-last network every Conv3x3 therefore costs a layout copy in front of the pad, the NCHW pad kernel, a copy
// back for the convolution, and the same again in backward (whose atomics-based kernel also leaves NCHW gradients
// that turn the following elu / add / upsample backward kernels into strided ones): 5 ms of a 19 ms training step
// (profiles/r2m_train_step_kernels.txt).  These kernels are pure data movement at HBM speed and keep NHWC tensors
// NHWC.  Values are bit-identical to ATen's forward; the backward adds the <= 9 contributions of an element in a
// fixed order (ATen: atomics).
//
// One CTA row per image row: blockIdx.x = n * rows + y, blockIdx.y * 256 + threadIdx.x = vector index inside the row
// (x * C/VEC + c).  NHWC with C % 4 == 0 moves float4; NCHW is the same kernel with C = 1 on N*C planes.
#include <cuda_runtime.h>

#include "../../include/md2_ops.h"
#include "md2_nvtx.h"

namespace md2 {

template <int VEC> struct Vec;
template <> struct Vec<1> { typedef float T; };
template <> struct Vec<4> { typedef float4 T; };

__device__ __forceinline__ float vadd4(float a, float b) { return a + b; }
__device__ __forceinline__ float4 vadd4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

struct PadShape {
  int C, H, W, Ho, Wo, pl, pt, pr, pb;
};

template <int VEC>
__global__ void __launch_bounds__(256) reflect_pad_fwd(const float* __restrict__ in, float* __restrict__ out,
                                                        const __grid_constant__ PadShape s) {
  typedef typename Vec<VEC>::T V;
  const int cv = s.C / VEC;
  const unsigned j = blockIdx.y * 256u + threadIdx.x;
  if (j >= (unsigned)(s.Wo * cv)) return;
  const int xo = j / cv, c = j - xo * cv;
  const unsigned row = blockIdx.x;  // n * Ho + yo
  const unsigned n = row / s.Ho;
  const int yo = row - n * s.Ho;
  const int y = reflect_index(yo - s.pt, s.H), x = reflect_index(xo - s.pl, s.W);
  const V* src = reinterpret_cast<const V*>(in) + (((size_t)n * s.H + y) * s.W + x) * cv + c;
  V* dst = reinterpret_cast<V*>(out) + ((size_t)row * s.Wo + xo) * cv + c;
  *dst = __ldg(src);
}

// positions of the padded axis that read input index i: itself, its mirror in the leading pad, its mirror in the
// trailing pad (pads are smaller than the axis, so there is at most one of each)
__device__ __forceinline__ int pad_sources(int i, int n, int lead, int trail, int (&o)[3]) {
  int k = 0;
  o[k++] = i + lead;
  if (i >= 1 && i <= lead) o[k++] = lead - i;
  if (i <= n - 2 && i >= n - 1 - trail) o[k++] = lead + 2 * (n - 1) - i;
  return k;
}

template <int VEC>
__global__ void __launch_bounds__(256) reflect_pad_bwd(const float* __restrict__ g_out, float* __restrict__ g_in,
                                                        const __grid_constant__ PadShape s) {
  typedef typename Vec<VEC>::T V;
  const int cv = s.C / VEC;
  const unsigned j = blockIdx.y * 256u + threadIdx.x;
  if (j >= (unsigned)(s.W * cv)) return;
  const int x = j / cv, c = j - x * cv;
  const unsigned row = blockIdx.x;  // n * H + y
  const unsigned n = row / s.H;
  const int y = row - n * s.H;
  int ys[3], xs[3];
  const int ny = pad_sources(y, s.H, s.pt, s.pb, ys), nx = pad_sources(x, s.W, s.pl, s.pr, xs);
  const V* g = reinterpret_cast<const V*>(g_out) + (size_t)n * s.Ho * s.Wo * cv + c;
  V acc = __ldg(g + ((size_t)ys[0] * s.Wo + xs[0]) * cv);
  for (int a = 0; a < ny; ++a)
    for (int b = (a == 0 ? 1 : 0); b < nx; ++b) acc = vadd4(acc, __ldg(g + ((size_t)ys[a] * s.Wo + xs[b]) * cv));
  reinterpret_cast<V*>(g_in)[((size_t)row * s.W + x) * cv + c] = acc;
}

static int check_pad(int N, int C, int H, int W, int pl, int pr, int pt, int pb, PadShape* s, long long* rows_in,
                     long long* rows_out, int channels_last) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || pl < 0 || pr < 0 || pt < 0 || pb < 0) return MD2_ERR_SHAPE;
  // torch: "Padding size should be less than the corresponding input dimension"
  if (pl >= W || pr >= W || pt >= H || pb >= H) return MD2_ERR_SHAPE;
  const long long planes = channels_last ? N : (long long)N * C;
  s->C = channels_last ? C : 1;
  s->H = H; s->W = W; s->Ho = H + pt + pb; s->Wo = W + pl + pr;
  s->pl = pl; s->pt = pt; s->pr = pr; s->pb = pb;
  *rows_in = planes * H;
  *rows_out = planes * s->Ho;
  if (*rows_out > 0x7fffffffLL || (long long)s->Wo * s->C > 0x7fffffffLL) return MD2_ERR_SHAPE;
  return 0;
}

static inline bool aligned16(const void* a, const void* b) { return (((uintptr_t)a | (uintptr_t)b) & 15) == 0; }

}  // namespace md2

using namespace md2;

extern "C" {

int md2_reflection_pad2d_forward(int N, int C, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                                 int channels_last, const float* in, float* out, md2_stream_t stream) {
  PadShape s;
  long long rows_in, rows_out;
  const int v = check_pad(N, C, H, W, pad_l, pad_r, pad_t, pad_b, &s, &rows_in, &rows_out, channels_last);
  if (v != 0) return v;
  if (!in || !out) return MD2_ERR_NULL;
  const NvtxRange range("md2_reflection_pad2d_forward");
  const bool vec = (s.C % 4 == 0) && aligned16(in, out);
  const int row_vectors = s.Wo * (vec ? s.C / 4 : s.C);
  const dim3 grid((unsigned)rows_out, (unsigned)((row_vectors + 255) / 256));
  if (grid.y > 65535) return MD2_ERR_SHAPE;
  if (vec) reflect_pad_fwd<4><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, s);
  else reflect_pad_fwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, s);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int md2_reflection_pad2d_backward(int N, int C, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                                  int channels_last, const float* g_out, float* g_in, md2_stream_t stream) {
  PadShape s;
  long long rows_in, rows_out;
  const int v = check_pad(N, C, H, W, pad_l, pad_r, pad_t, pad_b, &s, &rows_in, &rows_out, channels_last);
  if (v != 0) return v;
  if (!g_out || !g_in) return MD2_ERR_NULL;
  const NvtxRange range("md2_reflection_pad2d_backward");
  const bool vec = (s.C % 4 == 0) && aligned16(g_out, g_in);
  const int row_vectors = s.W * (vec ? s.C / 4 : s.C);
  const dim3 grid((unsigned)rows_in, (unsigned)((row_vectors + 255) / 256));
  if (grid.y > 65535) return MD2_ERR_SHAPE;
  if (vec) reflect_pad_bwd<4><<<grid, 256, 0, (cudaStream_t)stream>>>(g_out, g_in, s);
  else reflect_pad_bwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g_out, g_in, s);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
