// md2_host.h - host-side logic shared by the CUDA library (md2_abi.cu) and the host
// emulation used by the CPU-only tests: argument validation, the kernel parameter block
// and the workspace layout.  Plain C++, no CUDA, no torch.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/md2_loss.h"
#include "md2_tile.cuh"

namespace md2 {

// tile shapes of the two kernel families (see DESIGN.md "Data layout")
#ifndef MD2_TW
#define MD2_TW 32
#endif
#ifndef MD2_TH
#define MD2_TH 16
#endif
#ifndef MD2_NT
#define MD2_NT 256
#endif
#ifndef MD2_MINB
#define MD2_MINB 2
#endif
constexpr int kTW = MD2_TW;
constexpr int kTH = MD2_TH;
constexpr int kNT = MD2_NT;
// Tile height per source count: the backward build keeps S-proportional buffers in shared memory, so the tile
// shrinks with S to keep two CTAs resident per SM (<= ~113 KB each; DESIGN.md 4 lists the sizes).
#ifndef MD2_TH3
#define MD2_TH3 12
#endif
#ifndef MD2_TH4
#define MD2_TH4 12
#endif
constexpr int tile_h(int S) { return S <= 2 ? kTH : (S == 3 ? MD2_TH3 : MD2_TH4); }
// Threads per CTA of the training build, chosen so that the cells of phase A (36 x (TH+4)), the two-window runs of
// phase B (34 x (TH+2)/2) and the pixel runs of phase C (32 x TH/2) fill whole passes: 320 threads for the 32x16
// tile (720 cells / 306 runs / 256 runs); the 32x12 tile of S >= 3 (576 / 238 / 192) needs 128 registers per thread
// for its two units and keeps 256 (288 threads at 112 registers spill: 1.83 vs 1.16 ms at S = 3).  The forward-only
// build has 32 x TH/2 runs and uses kNT.
#ifndef MD2_NT_BWD
#define MD2_NT_BWD 320
#endif
#ifndef MD2_NT_BWD34
#define MD2_NT_BWD34 256
#endif
constexpr int tile_nt(int S, bool bwd) { return !bwd ? kNT : (S <= 2 ? MD2_NT_BWD : MD2_NT_BWD34); }

enum Mode { kForward = 0, kFused = 1, kBackward = 2 };

// Which rounding torch.matmul applies to the per-pixel dot products (see Tile's MM_ parameter)
inline int matmul_mode(int B, int H, int W) {
  if (B > 1) return 0;
  const long long n = (long long)H * W;
  if (3 * n >= 786432) return 0;
  return 4 * n >= 786432 ? 2 : 1;
}

inline int validate_cfg(const md2_cfg* c) {
  if (!c) return MD2_ERR_NULL;
  if (c->B < 1 || c->H < 4 || c->W < 4) return MD2_ERR_SHAPE;
  if (c->S < 1 || c->S > MD2_MAX_SOURCES) return MD2_ERR_SHAPE;
  if (c->num_scales < 1 || c->num_scales > MD2_MAX_SCALES) return MD2_ERR_SHAPE;
  const int div = 1 << (c->num_scales - 1);
  if (c->H % div != 0 || c->W % div != 0) return MD2_ERR_SHAPE;
  if ((c->H >> (c->num_scales - 1)) < 2 || (c->W >> (c->num_scales - 1)) < 2) return MD2_ERR_SHAPE;
  if ((int64_t)c->B * c->H * c->W * 4 > (int64_t)0x7fffffff) return MD2_ERR_SHAPE;
  if (!(c->min_depth > 0.0) || !(c->max_depth > c->min_depth)) return MD2_ERR_CONFIG;
  return 0;
}

inline int validate_inputs(const md2_cfg* c, const md2_inputs* in) {
  if (!in) return MD2_ERR_NULL;
  if (!in->target || !in->K || !in->inv_K) return MD2_ERR_NULL;
  for (int f = 0; f < c->S; ++f)
    if (!in->sources[f] || !in->T[f]) return MD2_ERR_NULL;
  for (int s = 0; s < c->num_scales; ++s)
    if (!in->disp[s] || !in->color_pyr[s]) return MD2_ERR_NULL;
  return 0;
}

struct Workspace {
  size_t off_tile_loss, off_dP, off_smooth, bytes;
  int tiles_x, tiles_y, n_tiles;
};

inline Workspace workspace_layout(const md2_cfg* c) {
  Workspace w;
  w.tiles_x = (c->W + kTW - 1) / kTW;
  w.tiles_y = (c->H + tile_h(c->S) - 1) / tile_h(c->S);
  w.n_tiles = w.tiles_x * w.tiles_y * c->B;
  size_t o = 0;
  w.off_tile_loss = o;
  o += (size_t)w.n_tiles * kMaxScales * sizeof(float);
  w.off_dP = o;
  o += (size_t)w.n_tiles * c->S * 12 * sizeof(float);
  w.off_smooth = o;
  o += (size_t)c->B * smooth_total(c->num_scales) * 3 * sizeof(float);
  w.bytes = (o + 255) & ~(size_t)255;
  return w;
}

inline void fill_params(Params& p, const md2_cfg* c, const md2_inputs* in, const md2_outputs* out,
                        const md2_grads* g, void* workspace, Mode mode) {
  memset(&p, 0, sizeof(p));
  p.B = c->B; p.H = c->H; p.W = c->W; p.S = c->S; p.ns = c->num_scales;
  p.automask = c->automask ? 1 : 0;
  p.use_saved_k = mode == kBackward;
  p.kt_fma = c->B > 1;
  // disparity2depth evaluates its scalars in Python doubles (warp.py:34-37)
  const double min_disp = 1.0 / c->max_depth, max_disp = 1.0 / c->min_depth;
  p.a = (float)min_disp;
  p.r = (float)(max_disp - min_disp);
  p.eps = (float)c->eps_proj;
  p.wm1 = (float)(c->W - 1);
  p.hm1 = (float)(c->H - 1);
  p.inv_wm1 = 1.0f / p.wm1;  // ATen CUDA: tensor / python_scalar == tensor * (1.0f / scalar)
  p.inv_hm1 = 1.0f / p.hm1;
  p.c1 = (float)0.0001;  // C1 = 0.01 ** 2 evaluates to the double 0.0001 in Python (model_loss.py:25)
  p.c2 = (float)0.0009;  // C2 = 0.03 ** 2 evaluates to the double 0.0009            (model_loss.py:26)
  p.lambda = (float)c->disp_smoothness;
  p.target = in->target;
  for (int f = 0; f < c->S; ++f) { p.src[f] = in->sources[f]; p.T[f] = in->T[f]; }
  for (int s = 0; s < c->num_scales; ++s) {
    p.disp[s] = in->disp[s];
    p.color[s] = in->color_pyr[s];
    p.noise[s] = in->noise[s];
    if (g) p.grad_disp[s] = g->grad_disp[s];
  }
  p.K = in->K; p.invK = in->inv_K; p.seed = in->seed;
  mix_seed(in->seed, p.seed_m1, p.seed_m2);
  p.seed_dev = (const unsigned long long*)in->seed_dev;
  if (out) { p.per_px = out->per_pixel; p.argmin = out->argmin; p.depth = out->depth; }
  const Workspace w = workspace_layout(c);
  char* ws = (char*)workspace;
  p.tile_loss = (float*)(ws + w.off_tile_loss);
  p.dP_part = (float*)(ws + w.off_dP);
  p.smooth_part = (float*)(ws + w.off_smooth);
  p.tiles_x = w.tiles_x; p.tiles_y = w.tiles_y; p.n_tiles = w.n_tiles;
  p.gcoef = (float)(1.0 / ((double)c->num_scales * c->B * c->H * c->W));
  // scale-0 gradients leave the tile as 16-byte vector reductions when the rows allow it
  p.vec_atomics = (g && g->grad_disp[0] && ((uintptr_t)g->grad_disp[0] & 15) == 0 && (c->W & 3) == 0) ? 1 : 0;
  p.grad_loss_host = 1.0f;
  p.dbg_scale = -1; p.dbg_source = -1;
}

}  // namespace md2
