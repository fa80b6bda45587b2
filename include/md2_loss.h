/*
 * md2_loss.h - C ABI of the B200-native view-synthesis loss (sm_100a).
 *
 * The reference (russellgeum/Digging-into-Self-Supervised-Monocular-Depth-Estimation)
 * is pure Python and has no FFI; its boundary for this path is the Python symbol
 * surface.  This header is the torch-free layer underneath the
 * PyTorch C++ extension: plain device pointers and sizes, no torch types.  Each entry
 * point names the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer to fp32, contiguous NCHW, unless noted;
 *   - return value 0 = success, <0 = invalid argument (MD2_ERR_*), >0 = cudaError_t of
 *     the launch; the functions never throw, never synchronise, never allocate: the
 *     caller owns all memory including the workspace (md2_workspace_bytes);
 *   - thread-safe for distinct streams with distinct workspaces.
 *
 * One "step" of the hot path = md2_loss_forward_backward (training) or
 * md2_loss_forward (validation under no_grad).
 */
#ifndef MD2_LOSS_H_
#define MD2_LOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MD2_MAX_SOURCES 4
#define MD2_MAX_SCALES 4

#define MD2_ERR_NULL      (-1)  /* a required pointer is NULL                         */
#define MD2_ERR_SHAPE     (-2)  /* B/H/W/S/num_scales out of range or not divisible    */
#define MD2_ERR_CONFIG    (-3)  /* inconsistent flags                                  */
#define MD2_ERR_WORKSPACE (-4)  /* workspace missing                                   */
#define MD2_ERR_NO_DEVICE (-5)  /* no CUDA device / wrong architecture                 */

typedef void* md2_stream_t; /* a cudaStream_t */

/* opt fields read by model_tool/processor.py:140-216 (model_option.py:5-89 defaults) */
typedef struct md2_cfg {
  int B, H, W;            /* batch, full-resolution height / width (H, W multiples of 2^(num_scales-1)) */
  int S;                  /* number of source frames = len(frame_ids) - 1, 1..MD2_MAX_SOURCES          */
  int num_scales;         /* len(opt.scales), scale s has size (H>>s, W>>s), 1..MD2_MAX_SCALES          */
  int automask;           /* opt.use_automasking                                                      */
  double min_depth;       /* opt.min_depth (0.1)                                                      */
  double max_depth;       /* opt.max_depth (100)                                                      */
  double disp_smoothness; /* opt.disp_smoothness (1e-3)                                               */
  double eps_proj;        /* PointCloud2Pixel eps (1e-7), model_layer/warp.py:251                      */
} md2_cfg;

/* Inputs of one step.  Replaces the dict reads of compute.image2warping /
 * compute.compute_loss (model_tool/processor.py:139-218). */
typedef struct md2_inputs {
  const float* target;                   /* inputs[("color",0,0)]            [B,3,H,W]            */
  const float* sources[MD2_MAX_SOURCES]; /* inputs[("color",f,0)]            S x [B,3,H,W]        */
  const float* disp[MD2_MAX_SCALES];     /* outputs[("disp",s)]              [B,1,H>>s,W>>s]      */
  const float* color_pyr[MD2_MAX_SCALES];/* inputs[("color",0,s)]            [B,3,H>>s,W>>s]      */
  const float* K;                        /* inputs[("K",0)]                  [B,4,4]              */
  const float* inv_K;                    /* inputs[("inv_K",0)]              [B,4,4]              */
  const float* T[MD2_MAX_SOURCES];       /* outputs[("c2c",f,0)] / inputs["stereo"]  S x [B,4,4]  */
  const float* noise[MD2_MAX_SCALES];    /* N(0,1) draws of processor.py:195, [B,S,H,W] per scale;
                                            NULL => generated on the device from `seed`          */
  uint64_t seed;
  const uint64_t* seed_dev;              /* optional DEVICE pointer: when non-NULL the seed is read from it by the
                                            kernel instead of `seed` (a CUDA graph replays frozen parameters, so a
                                            captured training step advances this word itself)               */
} md2_inputs;

/* Outputs of the forward part. */
typedef struct md2_outputs {
  float* loss;         /* outputs["loss"], 1 float                                            (required) */
  float* per_pixel;    /* min-reprojection value per pixel [num_scales,B,H,W]                 (optional) */
  uint8_t* argmin;     /* index into cat(identity, reprojection) [num_scales,B,H,W]           (optional;
                          required by md2_loss_backward)                                                */
  float* depth;        /* outputs[("depth",0,s)] [num_scales,B,1,H,W]                         (optional) */
} md2_outputs;

/* Gradients.  grad_disp[s] has the shape of disp[s]; grad_T[f] is [B,4,4] (pass NULL for a
 * source whose T is data, e.g. the stereo baseline).  All are overwritten, not accumulated. */
typedef struct md2_grads {
  float* grad_disp[MD2_MAX_SCALES];
  float* grad_T[MD2_MAX_SOURCES];
} md2_grads;

/* Bytes of scratch the three step functions need for this cfg (0 on invalid cfg). */
size_t md2_workspace_bytes(const md2_cfg* cfg);

/* Forward only (validation, model_train.py:76-79): image2warping + compute_loss. */
int md2_loss_forward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                     void* workspace, md2_stream_t stream);

/* Training step: forward and backward in one fused pass.  Gradients are those of
 * `grad_loss * loss` (grad_loss is a host scalar, 1.0 for loss.backward()). Replaces
 * image2warping + compute_loss + the autograd graph of model_train.py:68. */
int md2_loss_forward_backward(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                              const md2_grads* grads, float grad_loss,
                              void* workspace, md2_stream_t stream);

/* Stand-alone backward: recomputes the warps and window statistics from the inputs and the
 * argmin saved by md2_loss_forward (nothing else is stored between the two calls).
 * `grad_loss_dev` is a DEVICE pointer to the upstream scalar gradient. */
int md2_loss_backward(const md2_cfg* cfg, const md2_inputs* in, const uint8_t* argmin,
                      const float* grad_loss_dev, const md2_grads* grads,
                      void* workspace, md2_stream_t stream);

/* Pose parametrisation (model_layer/warp.py:43-153, param2matrix): axis-angle +
 * translation [n,3] each -> [n,4,4]; invert != 0 gives R^T * T(-t).  The backward takes
 * dL/dM [n,4,4] and returns dL/d axisangle and dL/d translation [n,3]. */
int md2_pose_forward(int n, const float* axisangle, const float* translation, int invert,
                     float* M, md2_stream_t stream);
int md2_pose_backward(int n, const float* axisangle, const float* translation, int invert,
                      const float* grad_M, float* grad_axisangle, float* grad_translation,
                      md2_stream_t stream);

/* Number of kernels the last call of each step function enqueued (for bench.py's
 * gpu_launches claim) and a version string "md2loss <x.y> sm_100a". */
int md2_launches_per_step(const md2_cfg* cfg, int with_backward);
const char* md2_version(void);

/* md2_loss_forward_backward with a per-call measurement hook (bench.py's roofline leg): records the two
 * caller-owned cudaEvent_t `start_event` / `stop_event` on `stream` immediately around the tile kernel - the
 * dominant launch - so that its duration can be read without a profiler.  No process-global state. */
int md2_loss_forward_backward_timed(const md2_cfg* cfg, const md2_inputs* in, const md2_outputs* out,
                                    const md2_grads* grads, float grad_loss, void* workspace,
                                    md2_stream_t stream, void* start_event, void* stop_event);

/* Debug taps used by the parity tests only: the sampling coordinates (ix, iy) in pixels
 * after un-normalisation and before clipping, and the warped image, for source f at
 * scale s.  coords [B,2,H,W], warped [B,3,H,W]. */
int md2_debug_warp(const md2_cfg* cfg, const md2_inputs* in, int scale, int source,
                   float* coords, float* warped, void* workspace, md2_stream_t stream);

/* Test hook: q_div[i] = num[i] / den[i] and q9[i] = num[i] / 9 computed with the kernels' guard-free
 * division sequences (csrc/md2_tile.cuh div_pos / div9), to be compared with IEEE division. */
int md2_debug_div(int n, const float* num, const float* den, float* q_div, float* q9, md2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MD2_LOSS_H_ */
