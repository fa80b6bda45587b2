// md2_emu.cpp - HOST EMULATION of the CUDA kernels (test infrastructure only).
//
// Compiles the phase functions of csrc/md2_tile.cuh with g++ and runs the threads of each
// CTA as a loop between barriers.  It lets the CPU-only test-suite check the kernel logic
// (indexing, halos, reflection, the analytic backward) against the oracle without a GPU.
// The product never links or loads this file.
#include <stdlib.h>
#include <vector>

#include "md2_host.h"

using namespace md2;

#define PHASE(stmt) \
  for (int tid = 0; tid < TK::NT; ++tid) { stmt; }

template <class TK>
static void run_tiles(const Params& p) {
  // the host build keeps one reduction row per thread, which may exceed the aliased warped-tile buffer:
  // give the arena enough room behind OFF_RED (everything after it is dead in the epilogue)
  std::vector<float> sm(TK::OFF_RED + TK::NT * TK::NRED + TK::SMEM_FLOATS + 64, 0.f);
  std::vector<typename TK::Regs> regs(TK::NT);
  for (int tile = 0; tile < p.n_tiles; ++tile) {
    typename TK::Ctx c;
    const int per_img = p.tiles_x * p.tiles_y, t = tile % per_img;
    TK::make_ctx(c, p, sm.data(), t % p.tiles_x, t / p.tiles_x, tile / per_img);
    PHASE(TK::init_regs(regs[tid]));
    PHASE(TK::setup(c, tid));
    if (p.use_tma) {
      PHASE(TK::load_tiles_zero_fill(c, tid));
      PHASE(TK::load_sources(c, tid));
      PHASE(TK::patch_border(c, tid));
    } else {
      PHASE(TK::load_tiles(c, tid));
    }
    PHASE(TK::prologue_windows(c, tid));
    for (int s = 0; s < p.ns; ++s) {
      if (TK::BWD && s > 0) PHASE(TK::phase_d2(c, s - 1, tid));
      PHASE(TK::template phase_a<true>(c, s, tid));
      PHASE(TK::phase_b(c, s, tid, regs[tid]));
      if (TK::BWD) {
        PHASE(TK::phase_c(c, s, tid, regs[tid]));
        PHASE(TK::phase_d1(c, s, tid));
      }
    }
    if (TK::BWD) PHASE(TK::phase_d2(c, p.ns - 1, tid));
    PHASE(TK::epilogue1(c, tid, regs[tid]));
    PHASE(TK::epilogue2(c, tid));
  }
}

template <bool BWD, int MM>
static void dispatch_tiles_s(const Params& p) {
  switch (p.S) {
    case 1: run_tiles<Tile<1, BWD, kTW, tile_h(1), tile_nt(1, BWD), MM>>(p); break;
    case 2: run_tiles<Tile<2, BWD, kTW, tile_h(2), tile_nt(2, BWD), MM>>(p); break;
    case 3: run_tiles<Tile<3, BWD, kTW, tile_h(3), tile_nt(3, BWD), MM>>(p); break;
    default: run_tiles<Tile<4, BWD, kTW, tile_h(4), tile_nt(4, BWD), MM>>(p); break;
  }
}
template <bool BWD>
static void dispatch_tiles(const Params& p) {
  switch (matmul_mode(p.B, p.H, p.W)) {
    case 0: dispatch_tiles_s<BWD, 0>(p); break;
    case 1: dispatch_tiles_s<BWD, 1>(p); break;
    default: dispatch_tiles_s<BWD, 2>(p); break;
  }
}

static void run_smooth_forward(const Params& p, bool zero_grad) {
  const int nt = 256;
  std::vector<float> red(nt * 3);
  for (int blk = 0; blk < p.B * smooth_total(p.ns); ++blk) {
    const SmoothBand k = smooth_band(p, blk);
    for (int tid = 0; tid < nt; ++tid) {
      float v[3];
      smooth_fwd_thread(p, k, tid, nt, zero_grad, v);
      Reduce<256>::stage1(v, 3, tid, red.data());
    }
    for (int i = 0; i < 3; ++i) p.smooth_part[(size_t)blk * 3 + i] = Reduce<256>::stage2(i, 3, red.data());
  }
}

static void run_smooth_backward(const Params& p) {
  const int nt = 256;
  const float gl = p.grad_loss_dev ? *p.grad_loss_dev : p.grad_loss_host;
  for (int blk = 0; blk < p.B * smooth_total(p.ns); ++blk) {
    const SmoothBand k = smooth_band(p, blk);
    const SmoothStats st = smooth_stats(p, k.s, k.b);
    for (int tid = 0; tid < nt; ++tid) smooth_bwd_thread(p, k, st, tid, nt, gl);
  }
}

static void run_finalize(const Params& p, float* loss, float* const* grad_T) {
  if (loss) {
    double acc = 0.0;
    for (int tid = 0; tid < 256; ++tid) acc += finalize_loss_partial(p, tid, 256);
    *loss = (float)acc;
  }
  if (grad_T)
    for (int idx = 0; idx < p.S * p.B * 16; ++idx) finalize_grad_T(p, grad_T, idx);
}

static int run(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out, const md2_grads* g,
               float grad_loss, const uint8_t* saved_k, Mode mode, Params* tweak = nullptr) {
  int e = validate_cfg(cfg);
  if (e) return e;
  e = validate_inputs(cfg, in);
  if (e) return e;
  if (mode != kBackward && (!out || !out->loss)) return MD2_ERR_NULL;
  if (mode != kForward) {
    if (!g) return MD2_ERR_NULL;
    for (int s = 0; s < cfg->num_scales; ++s)
      if (!g->grad_disp[s]) return MD2_ERR_NULL;
  }
  if (mode == kBackward && !saved_k) return MD2_ERR_NULL;
  const Workspace w = workspace_layout(cfg);
  std::vector<char> ws(w.bytes, 0);
  Params p;
  fill_params(p, cfg, in, out, g, ws.data(), mode);
  p.grad_loss_host = grad_loss;
  p.saved_k = saved_k;
  p.use_tma = getenv("MD2_EMU_TMA") != nullptr;
  if (tweak) {
    p.dbg_coords = tweak->dbg_coords;
    p.dbg_warped = tweak->dbg_warped;
    p.dbg_scale = tweak->dbg_scale;
    p.dbg_source = tweak->dbg_source;
  }
  run_smooth_forward(p, mode != kForward);
  if (mode == kForward) {
    dispatch_tiles<false>(p);
  } else {
    dispatch_tiles<true>(p);
    run_smooth_backward(p);
  }
  run_finalize(p, mode == kBackward ? nullptr : out->loss, mode == kForward ? nullptr : g->grad_T);
  return 0;
}

extern "C" {

int md2_emu_forward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out) {
  return run(cfg, in, out, nullptr, 1.0f, nullptr, kForward);
}

int md2_emu_forward_backward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                             const md2_grads* g, float grad_loss) {
  return run(cfg, in, out, g, grad_loss, nullptr, kFused);
}

int md2_emu_backward(const md2_cfg* cfg, const md2_inputs* in, const uint8_t* argmin, float grad_loss,
                     const md2_grads* g) {
  return run(cfg, in, nullptr, g, grad_loss, argmin, kBackward);
}

int md2_emu_debug_warp(const md2_cfg* cfg, const md2_inputs* in, int scale, int source, float* coords,
                       float* warped) {
  Params t;
  memset(&t, 0, sizeof(t));
  t.dbg_coords = coords;
  t.dbg_warped = warped;
  t.dbg_scale = scale;
  t.dbg_source = source;
  float loss;
  md2_outputs out;
  memset(&out, 0, sizeof(out));
  out.loss = &loss;
  return run(cfg, in, &out, nullptr, 1.0f, nullptr, kForward, &t);
}

int md2_emu_pose_forward(int n, const float* aa, const float* tr, int invert, float* M) {
  for (int i = 0; i < n; ++i) pose_forward_one(aa + 3 * i, tr + 3 * i, invert, M + 16 * i);
  return 0;
}

int md2_emu_pose_backward(int n, const float* aa, const float* tr, int invert, const float* gM, float* gaa,
                          float* gtr) {
  for (int i = 0; i < n; ++i)
    pose_backward_one(aa + 3 * i, tr + 3 * i, invert, gM + 16 * i, gaa + 3 * i, gtr + 3 * i);
  return 0;
}

}  // extern "C"
