// md2_abi.cu - sm_100a kernels and the extern "C" entry points of include/md2_loss.h.
//
// Launch structure of one step (5 launches with backward, 3 without; no host synchronisation, no allocation;
// the smoothness kernels run on a side stream beside the tile kernel):
//   0. zero_grads_kernel       (backward) zero-fill of the gradient buffers;
//   1. smooth_forward_kernel   per (scale, image, row band): sums for the mean-normalised
//                              smoothness (model_loss.py:77-88,112-116);
//  1b. smooth_backward_kernel  (backward) gradient of the smoothness term;
//   2. tile_kernel<S, BWD>     one CTA per 32 x TH image tile, all scales and sources
//                              (md2_tile.cuh; 3 CTA barriers per scale);
//   3. finalize_kernel         fixed-order reduction of the per-CTA partials -> loss,
//                              dL/dT = K^T dL/dP.
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)
#include <cuda_runtime.h>

#include <stdlib.h>

#include "md2_host.h"
#include "md2_nvtx.h"

// phase staggering of the co-resident CTAs (see tile_kernel): 0 off, 1 the CTAs nsm..2*nsm-1, 2 the odd CTAs below 2*nsm
#ifndef MD2_STAGGER_MODE
#define MD2_STAGGER_MODE 0
#endif
#ifndef MD2_STAGGER_CYCLES
#define MD2_STAGGER_CYCLES 16000
#endif

// phase-skipping experiments only (tools/variants.py): bit 0 phase A, 1 phase B, 2 phase C, 3 phase D1 (row pass /
// scale-0 reductions), 4 phase D2 (column pass), 5 the dL/dP warp reduction inside D1
#ifndef MD2_SKIP
#define MD2_SKIP 0
#endif

namespace md2 {

// ---------------------------------------------------------------------------------------- TMA
// Tensor maps over the NCHW fp32 images: dims (W, H, 3, B), box (R2P, R2H, 3, 1) = one tile with its halo,
// all three channels, out-of-image elements zero-filled.  The box x-origin is a multiple of 4 pixels: TMA
// needs a 16-byte aligned start in global memory (an unaligned start raises an illegal-instruction error).
struct alignas(64) TmaMaps {
  CUtensorMap target;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lookup_encode_tiled() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    return (EncodeTiledFn)ptr;
  return nullptr;
}
static EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = lookup_encode_tiled();  // function-local static: initialised once, thread-safe
  return fn;
}

static bool make_image_map(CUtensorMap* m, const float* img, int B, int H, int W, int box_w, int box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || ((uintptr_t)img & 15) != 0 || (W & 3) != 0 || (box_w & 3) != 0 || box_w > 256 || box_h > 256) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 3, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_tile(float* dst, const CUtensorMap* map, int x, int y, int b, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_addr(dst)), "l"((uint64_t)map), "r"(x), "r"(y), "r"(0), "r"(b), "r"(mbar)
      : "memory");
}

// Stage the target tile (+ halo) with TMA.
template <class TK>
__device__ __forceinline__ void tma_stage_tiles(const typename TK::Ctx& c, const Params& p, const TmaMaps& maps, float* sm,
                                                int tid) {
  const uint32_t mbar = smem_addr(sm + TK::OFF_MBAR);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)(TK::R2S * 3 * sizeof(float));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    // the box starts XO pixels left of the halo so that its first element is 16-byte aligned in global memory
    tma_load_tile(sm + TK::OFF_T, &maps.target, c.tx0 - TK::HB - TK::XO, c.ty0 - TK::HB, c.b, mbar);
  }
  __syncthreads();  // the barrier is initialised (and armed) before anybody polls it
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  // bounded poll: a TMA fault must abort the kernel, never hang the GPU
#pragma unroll 1
  for (int it = 0; it < (1 << 24); ++it) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// grid (tiles_x, tiles_y, B): one CTA per image tile
// register budget: MD2_MINB CTAs of the training build per SM (64 K registers, allocated in units of 8 per thread)
template <class TK>
constexpr int tile_max_regs() {
  const int r = TK::BWD ? (65536 / (MD2_MINB * TK::NT)) / 8 * 8 : 255;
  return r > 255 ? 255 : r;
}
template <class TK, bool DBG>
__global__ void __launch_bounds__(TK::NT) __maxnreg__(tile_max_regs<TK>())
    tile_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps maps) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x;
  typename TK::Ctx c;
  TK::make_ctx(c, p, sm, blockIdx.x, blockIdx.y, blockIdx.z);
  typename TK::Regs regs;
  TK::init_regs(regs);
  if (TK::BWD && MD2_STAGGER_MODE != 0) {
    // Phase-staggering experiment (off by default; measured: no effect, profiles/r2_experiments.json).  If the two
    // CTAs that share an SM ran in lockstep - both in the fp32-bound phase B, then both in the latency-bound phases A
    // and C - delaying the second CTA of every SM in the FIRST wave by about half a scale iteration would put them
    // out of phase for the whole launch (their successors inherit the offset).
    const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    unsigned nsm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const bool second = MD2_STAGGER_MODE == 1 ? (lin >= nsm && lin < 2 * nsm) : (lin < 2 * nsm && (lin & 1));
    if (second) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)MD2_STAGGER_CYCLES) __nanosleep(256);
    }
  }
  if (p.use_tma) {
    // TMA: one thread issues the box loads, everybody computes K.T meanwhile, then waits on the mbarrier
    tma_stage_tiles<TK>(c, p, maps, sm, tid);
    TK::setup(c, tid);
    TK::load_sources(c, tid);  // raw sources for the identity loss, overlapping the TMA transfer
    mbar_wait(smem_addr(sm + TK::OFF_MBAR), 0);
    TK::patch_border(c, tid);  // reflection at the image border (TMA zero-fills)
  } else {
    TK::setup(c, tid);
    TK::load_tiles(c, tid);
  }
  __syncthreads();
  TK::prologue_windows(c, tid);
  __syncthreads();
  for (int s = 0; s < p.ns; ++s) {
    if (TK::BWD && s > 0 && !(MD2_SKIP & 16)) TK::phase_d2(c, s - 1, tid);  // column pass of the previous scale's upsample adjoint
    if (!(MD2_SKIP & 1)) TK::template phase_a<DBG>(c, s, tid);
    __syncthreads();
    if (!(MD2_SKIP & 2)) TK::phase_b(c, s, tid, regs);
    __syncthreads();
    if (TK::BWD) {
      if (!(MD2_SKIP & 4)) TK::phase_c(c, s, tid, regs);
      __syncwarp();  // phase D1 reads the two tile rows its own warp has just produced
      if (!(MD2_SKIP & 8)) TK::phase_d1(c, s, tid);
      __syncthreads();
    }
  }
  if (TK::BWD && !(MD2_SKIP & 16)) TK::phase_d2(c, p.ns - 1, tid);
  TK::epilogue1(c, tid, regs);
  __syncthreads();
  TK::epilogue2(c, tid);
}

__global__ void __launch_bounds__(256) smooth_forward_kernel(const __grid_constant__ Params p, int zero_grad) {
  __shared__ float red[8 * 3];
  const SmoothBand k = smooth_band(p, blockIdx.x);
  float v[3];
  smooth_fwd_thread(p, k, threadIdx.x, 256, false, v);
  (void)zero_grad;
  Reduce<256>::stage1(v, 3, threadIdx.x, red);
  __syncthreads();
  if (threadIdx.x < 3) p.smooth_part[(size_t)blockIdx.x * 3 + threadIdx.x] = Reduce<256>::stage2(threadIdx.x, 3, red);
}

// zero-fill of the gradient buffers (all scales), before anything accumulates into them
__global__ void __launch_bounds__(256) zero_grads_kernel(const __grid_constant__ Params p) {
  for (int s = 0; s < p.ns; ++s) {
    const int n = p.B * (p.H >> s) * (p.W >> s);
    float* g = p.grad_disp[s];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) g[i] = 0.f;
  }
}

// gradient of the smoothness term, one block per row band (atomics into grad_disp; independent of the
// tile kernel, so it runs beside it on the side stream)
__global__ void __launch_bounds__(256) smooth_backward_kernel(const __grid_constant__ Params p) {
  __shared__ float stat[3];
  const int t = threadIdx.x;
  const SmoothBand k = smooth_band(p, blockIdx.x);
  const float gl = p.grad_loss_dev ? __ldg(p.grad_loss_dev) : p.grad_loss_host;
  if (t < 3) {  // per-image statistics once per block (fixed-order sum over the row bands)
    const float* part = p.smooth_part + ((size_t)k.b * smooth_total(p.ns) + smooth_offset(k.s)) * 3;
    float acc = 0.f;
    for (int ch = 0; ch < smooth_chunks(k.s); ++ch) acc += part[ch * 3 + t];
    stat[t] = acc;
  }
  __syncthreads();
  SmoothStats st;
  st.inv = 1.0f / (stat[0] / (float)(k.hs * k.ws) + 1e-7f);
  st.sx = stat[1];
  st.sy = stat[2];
  smooth_bwd_thread(p, k, st, t, 256, gl);
}

struct GradTPtrs {
  float* p[kMaxS];
};

// block 0: the scalar loss; blocks 1 + (f * B + b): dL/dT_f[b] = K_b^T (sum over the image's tiles of dL/dP)
constexpr int kFinNT = 1024;  // block 0 is latency bound (dependent loads, fp64 divisions): 32 warps instead of 8
__global__ void __launch_bounds__(kFinNT) finalize_kernel(const __grid_constant__ Params p, float* loss, GradTPtrs gT,
                                                          int want_grad_T) {
  __shared__ double red[kFinNT];
  __shared__ float part[21 * 12];
  __shared__ float dPs[12];
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    if (!loss) return;
    // photometric part: per-CTA partial sums, thread-strided in double
    double acc = 0.0;
    const double inv_n = 1.0 / ((double)p.B * p.H * p.W);
    for (int i = t; i < p.n_tiles * kMaxScales; i += kFinNT)
      if ((i % kMaxScales) < p.ns) acc += (double)p.tile_loss[i] * inv_n;
    // smoothness part: one warp per (scale, image), lanes stride the row bands, fixed-order shuffle tree
    const int warp = t >> 5, lane = t & 31;
    for (int pair = warp; pair < p.ns * p.B; pair += kFinNT / 32) {
      const int s = pair / p.B, b = pair - s * p.B;
      const float* part = p.smooth_part + ((size_t)b * smooth_total(p.ns) + smooth_offset(s)) * 3;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int ch = lane; ch < smooth_chunks(s); ch += 32) {
        a0 += part[ch * 3 + 0];
        a1 += part[ch * 3 + 1];
        a2 += part[ch * 3 + 2];
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if (lane == 0) {
        const int hs = p.H >> s, ws = p.W >> s;
        const double inv = 1.0 / ((double)a0 / ((double)hs * ws) + 1e-7);
        double sm = 0.0;
        if (ws > 1) sm += inv * a1 / ((double)p.B * hs * (ws - 1));
        if (hs > 1) sm += inv * a2 / ((double)p.B * (hs - 1) * ws);
        acc += (double)p.lambda * sm / (double)(1 << s);
      }
    }
    red[t] = acc / (double)p.ns;
    __syncthreads();
    for (int o = kFinNT / 2; o > 0; o >>= 1) {  // fixed-shape tree: deterministic
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) *loss = (float)red[0];
    return;
  }
  if (!want_grad_T) return;
  const int fb = blockIdx.x - 1;
  const int b = fb % p.B, f = fb / p.B;
  if (f >= p.S || gT.p[f] == nullptr) return;
  const int per_img = p.tiles_x * p.tiles_y;
  if (t < 21 * 12) {
    const int v = t % 12, g = t / 12;
    float acc = 0.f;
    for (int tile = g; tile < per_img; tile += 21)
      acc += p.dP_part[((size_t)(b * per_img + tile) * p.S + f) * 12 + v];
    part[g * 12 + v] = acc;
  }
  __syncthreads();
  if (t < 12) {
    float acc = 0.f;
    for (int g = 0; g < 21; ++g) acc += part[g * 12 + t];
    dPs[t] = acc;
  }
  __syncthreads();
  if (t < 16) {
    const int k = t >> 2, j = t & 3;
    const float* K = p.K + b * 16;
    gT.p[f][b * 16 + t] = K[0 * 4 + k] * dPs[0 * 4 + j] + K[1 * 4 + k] * dPs[1 * 4 + j] + K[2 * 4 + k] * dPs[2 * 4 + j];
  }
}

__global__ void pose_forward_kernel(int n, const float* aa, const float* tr, int invert, float* M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pose_forward_one(aa + 3 * i, tr + 3 * i, invert, M + 16 * i);
}
__global__ void pose_backward_kernel(int n, const float* aa, const float* tr, int invert, const float* gM, float* gaa,
                                     float* gtr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pose_backward_one(aa + 3 * i, tr + 3 * i, invert, gM + 16 * i, gaa + 3 * i, gtr + 3 * i);
}

__global__ void debug_div_kernel(int n, const float* num, const float* den, float* q_div, float* q9) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float rinv;
    q_div[i] = div_pos(num[i], den[i], rinv);
    q9[i] = div9(num[i]);
  }
}

// environment switches for A/B measurements, read once (thread-safe function-local statics)
static bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}
static bool tma_enabled() {   // MD2_NO_TMA=1 disables the TMA path
  static const bool v = !env_flag("MD2_NO_TMA");
  return v;
}
static bool side_enabled() {  // MD2_NO_SIDE=1 keeps the smoothness kernels on the caller's stream
  static const bool v = !env_flag("MD2_NO_SIDE");
  return v;
}

// per-call measurement hook (md2_loss_forward_backward_timed): events recorded around the tile kernel
struct TileEvents {
  cudaEvent_t start, stop;
};

constexpr int kMaxDevices = 64;

// occupancy experiments only (tools/variants.py): extra dynamic shared memory per CTA
#ifndef MD2_EXTRA_SMEM
#define MD2_EXTRA_SMEM 0
#endif

template <class TK, bool DBG>
static cudaError_t launch_tiles_impl(const Params& p, cudaStream_t st, const TileEvents* ev) {
  // The dynamic shared-memory limit is a per-device (per-context) function attribute: set it once per device.
  // A racing second thread only repeats the (idempotent) call.
  static bool attr_set[kMaxDevices] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    e = cudaFuncSetAttribute(tile_kernel<TK, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)TK::SMEM_BYTES + MD2_EXTRA_SMEM);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  const dim3 grid(p.tiles_x, p.tiles_y, p.B);
  // TMA staging of the target / raw source tiles when the layout allows it (16-byte aligned bases, W % 4 == 0,
  // box row a multiple of 16 bytes); otherwise the kernel loads the tiles with plain coalesced loads
  TmaMaps maps;
  Params q = p;
  q.use_tma = tma_enabled() && (TK::TW % 4 == 0) && make_image_map(&maps.target, p.target, p.B, p.H, p.W, TK::R2P, TK::R2H);
  if (ev) cudaEventRecord(ev->start, st);
  tile_kernel<TK, DBG><<<grid, TK::NT, TK::SMEM_BYTES + MD2_EXTRA_SMEM, st>>>(q, maps);
  if (ev) cudaEventRecord(ev->stop, st);
  return cudaGetLastError();
}

template <class TK>
static cudaError_t launch_tiles(const Params& p, cudaStream_t st, const TileEvents* ev) {
  if (!TK::BWD && p.dbg_coords) return launch_tiles_impl<TK, !TK::BWD>(p, st, ev);  // debug tap: forward build only
  return launch_tiles_impl<TK, false>(p, st, ev);
}

template <bool BWD, int MM>
static cudaError_t dispatch_tiles_s(const Params& p, cudaStream_t st, const TileEvents* ev) {
  switch (p.S) {
    case 1: return launch_tiles<Tile<1, BWD, kTW, tile_h(1), tile_nt(1, BWD), MM>>(p, st, ev);
    case 2: return launch_tiles<Tile<2, BWD, kTW, tile_h(2), tile_nt(2, BWD), MM>>(p, st, ev);
    case 3: return launch_tiles<Tile<3, BWD, kTW, tile_h(3), tile_nt(3, BWD), MM>>(p, st, ev);
    default: return launch_tiles<Tile<4, BWD, kTW, tile_h(4), tile_nt(4, BWD), MM>>(p, st, ev);
  }
}
template <bool BWD>
static cudaError_t dispatch_tiles(const Params& p, cudaStream_t st, const TileEvents* ev) {
  switch (matmul_mode(p.B, p.H, p.W)) {
    case 0: return dispatch_tiles_s<BWD, 0>(p, st, ev);
    case 1: return dispatch_tiles_s<BWD, 1>(p, st, ev);
    default: return dispatch_tiles_s<BWD, 2>(p, st, ev);
  }
}

// One non-blocking side stream + fork/join events per (host thread, device); created on first use.
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
static SideStream* side_stream() {
  thread_local SideStream cache[kMaxDevices];
  thread_local bool ready[kMaxDevices] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  if (!ready[dev]) {
    SideStream s = {nullptr, nullptr, nullptr};
    const bool ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {  // release whatever was created: the next call starts from scratch
      if (s.join) cudaEventDestroy(s.join);
      if (s.fork) cudaEventDestroy(s.fork);
      if (s.stream) cudaStreamDestroy(s.stream);
      return nullptr;
    }
    cache[dev] = s;
    ready[dev] = true;
  }
  return &cache[dev];
}

static int run_step(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out, const md2_grads* g,
                    float grad_loss, const float* grad_loss_dev, const uint8_t* saved_k, void* workspace,
                    md2_stream_t stream, Mode mode, const Params* tweak, const TileEvents* ev = nullptr) {
  int e = validate_cfg(cfg);
  if (e) return e;
  e = validate_inputs(cfg, in);
  if (e) return e;
  if (!workspace) return MD2_ERR_WORKSPACE;
  if (mode != kBackward && (!out || !out->loss)) return MD2_ERR_NULL;
  if (mode != kForward) {
    if (!g) return MD2_ERR_NULL;
    for (int s = 0; s < cfg->num_scales; ++s)
      if (!g->grad_disp[s]) return MD2_ERR_NULL;
  }
  if (mode == kBackward && (!saved_k || !grad_loss_dev)) return MD2_ERR_NULL;
  Params p;
  fill_params(p, cfg, in, out, g, workspace, mode);
  p.grad_loss_host = grad_loss;
  p.grad_loss_dev = grad_loss_dev;
  p.saved_k = saved_k;
  if (tweak) {
    p.dbg_coords = tweak->dbg_coords;
    p.dbg_warped = tweak->dbg_warped;
    p.dbg_scale = tweak->dbg_scale;
    p.dbg_source = tweak->dbg_source;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // The smoothness term (forward sums, then its gradient) does not depend on the tile kernel, so it runs on a
  // side stream beside it (fork / join with events: no host synchronisation, capturable in a CUDA graph).
  SideStream* side = side_stream();
  if (!side) return MD2_ERR_NO_DEVICE;
  cudaError_t ce;
  if (mode != kForward) {
    zero_grads_kernel<<<296, 256, 0, st>>>(p);
    if ((ce = cudaGetLastError()) != cudaSuccess) return (int)ce;
  }
  const bool use_side = side_enabled();
  cudaStream_t ss = use_side ? side->stream : st;
  if (use_side) {
    if ((ce = cudaEventRecord(side->fork, st)) != cudaSuccess) return (int)ce;
    if ((ce = cudaStreamWaitEvent(ss, side->fork, 0)) != cudaSuccess) return (int)ce;
  }
  smooth_forward_kernel<<<p.B * smooth_total(p.ns), 256, 0, ss>>>(p, 0);
  if ((ce = cudaGetLastError()) != cudaSuccess) return (int)ce;
  if (mode != kForward) {
    smooth_backward_kernel<<<p.B * smooth_total(p.ns), 256, 0, ss>>>(p);
    if ((ce = cudaGetLastError()) != cudaSuccess) return (int)ce;
  }
  if (use_side && (ce = cudaEventRecord(side->join, ss)) != cudaSuccess) return (int)ce;
  ce = mode == kForward ? dispatch_tiles<false>(p, st, ev) : dispatch_tiles<true>(p, st, ev);
  if (ce != cudaSuccess) return (int)ce;
  if (use_side && (ce = cudaStreamWaitEvent(st, side->join, 0)) != cudaSuccess) return (int)ce;
  GradTPtrs gT;
  for (int f = 0; f < kMaxS; ++f) gT.p[f] = (g && f < cfg->S) ? g->grad_T[f] : nullptr;
  finalize_kernel<<<1 + (mode != kForward ? p.B * p.S : 0), kFinNT, 0, st>>>(p, mode == kBackward ? nullptr : out->loss, gT,
                                                                            mode != kForward);
  ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

}  // namespace md2

using namespace md2;

extern "C" {

size_t md2_workspace_bytes(const md2_cfg* cfg) {
  if (validate_cfg(cfg)) return 0;
  return workspace_layout(cfg).bytes;
}

int md2_loss_forward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out, void* workspace,
                     md2_stream_t stream) {
  const NvtxRange range("md2_loss_forward");
  return run_step(cfg, in, out, nullptr, 1.0f, nullptr, nullptr, workspace, stream, kForward, nullptr);
}

int md2_loss_forward_backward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                              const md2_grads* grads, float grad_loss, void* workspace, md2_stream_t stream) {
  const NvtxRange range("md2_loss_forward_backward");
  return run_step(cfg, in, out, grads, grad_loss, nullptr, nullptr, workspace, stream, kFused, nullptr);
}

int md2_loss_backward(const md2_cfg* cfg, const md2_inputs* in, const uint8_t* argmin, const float* grad_loss_dev,
                      const md2_grads* grads, void* workspace, md2_stream_t stream) {
  const NvtxRange range("md2_loss_backward");
  return run_step(cfg, in, nullptr, grads, 1.0f, grad_loss_dev, argmin, workspace, stream, kBackward, nullptr);
}

int md2_debug_warp(const md2_cfg* cfg, const md2_inputs* in, int scale, int source, float* coords, float* warped,
                   void* workspace, md2_stream_t stream) {
  if (!coords || !warped || !workspace) return MD2_ERR_NULL;
  Params t;
  memset(&t, 0, sizeof(t));
  t.dbg_coords = coords;
  t.dbg_warped = warped;
  t.dbg_scale = scale;
  t.dbg_source = source;
  // the scalar loss lands in the last 4 bytes of the (256-byte padded) workspace
  md2_outputs out;
  memset(&out, 0, sizeof(out));
  const size_t wb = md2_workspace_bytes(cfg);
  if (wb == 0) return MD2_ERR_SHAPE;
  out.loss = (float*)((char*)workspace + wb - sizeof(float));
  return run_step(cfg, in, &out, nullptr, 1.0f, nullptr, nullptr, workspace, stream, kForward, &t);
}

int md2_pose_forward(int n, const float* axisangle, const float* translation, int invert, float* M,
                     md2_stream_t stream) {
  if (n < 0) return MD2_ERR_SHAPE;
  if (!axisangle || !translation || !M) return MD2_ERR_NULL;
  if (n == 0) return 0;
  pose_forward_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n, axisangle, translation, invert, M);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int md2_pose_backward(int n, const float* axisangle, const float* translation, int invert, const float* grad_M,
                      float* grad_axisangle, float* grad_translation, md2_stream_t stream) {
  if (n < 0) return MD2_ERR_SHAPE;
  if (!axisangle || !translation || !grad_M || !grad_axisangle || !grad_translation) return MD2_ERR_NULL;
  if (n == 0) return 0;
  pose_backward_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n, axisangle, translation, invert, grad_M,
                                                                       grad_axisangle, grad_translation);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int md2_launches_per_step(const md2_cfg* cfg, int with_backward) {
  if (validate_cfg(cfg)) return 0;
  return with_backward ? 5 : 3;  // [zero grads,] smoothness sums, [smoothness gradient,] tile kernel, finalize
}

int md2_debug_div(int n, const float* num, const float* den, float* q_div, float* q9, md2_stream_t stream) {
  if (n <= 0) return MD2_ERR_SHAPE;
  if (!num || !den || !q_div || !q9) return MD2_ERR_NULL;
  md2::debug_div_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, num, den, q_div, q9);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int md2_loss_forward_backward_timed(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                                    const md2_grads* grads, float grad_loss, void* workspace, md2_stream_t stream,
                                    void* start_event, void* stop_event) {
  if (!start_event || !stop_event) return MD2_ERR_NULL;
  const TileEvents ev = {(cudaEvent_t)start_event, (cudaEvent_t)stop_event};
  const NvtxRange range("md2_loss_forward_backward");
  return run_step(cfg, in, out, grads, grad_loss, nullptr, nullptr, workspace, stream, kFused, nullptr, &ev);
}

const char* md2_version(void) { return "md2loss 0.1 sm_100a"; }

}  // extern "C"
