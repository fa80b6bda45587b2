/* md2_metrics.h - C ABI of the training-time depth metrics, SURVEY.md 8f row N2.
 *
 * Replaces compute_depth_metric + compute_depth_error(lib="torch") of model_loss/model_metric.py:52-106,
 * the other consumer of the fused loss's ("depth", 0, 0) output (model_tool/logger.py:30-36):
 *   bilinear up-sampling of the predicted depth to the ground-truth size (align_corners=False), clamp to
 *   [1e-3, 80], mask = gt > 0 inside the crop rows [y0, y1) x columns [x0, x1), median scaling
 *   pred *= median(gt) / median(pred) (torch.median: the LOWER median), clamp again, then
 *   abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 over the masked pixels of the whole batch.
 * The reference needs ~40 ATen launches, two sorts and 8 host syncs for this; here it is 9 launches, exact
 * medians by a three-pass radix select on the fp32 bit patterns, no sync, no allocation.
 *
 * Conventions as in md2_loss.h: device pointers to contiguous fp32, caller's stream, returns 0 /
 * MD2_ERR_* / cudaError_t.  out[0..6] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 (device, fp32);
 * out[7] = number of masked pixels (as a float).  With no masked pixel the seven metrics are NaN
 * (torch.median of an empty tensor raises in the reference).
 */
#ifndef MD2_METRICS_H_
#define MD2_METRICS_H_

#include "md2_loss.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct md2_metrics_cfg {
  int B, H, W;        /* predicted depth [B,1,H,W]                                  */
  int Hg, Wg;         /* ground truth    [B,1,Hg,Wg]     (375 x 1242 in the reference) */
  int y0, y1, x0, x1; /* crop, half-open (153, 371, 44, 1197 in the reference)      */
  float min_depth, max_depth; /* clamp range (1e-3, 80)                              */
} md2_metrics_cfg;

/* Bytes of scratch md2_depth_metrics needs (0 on an invalid cfg). */
size_t md2_metrics_workspace_bytes(const md2_metrics_cfg* cfg);

int md2_depth_metrics(const md2_metrics_cfg* cfg, const float* depth, const float* gt, float* out /* [8] */,
                      void* workspace, md2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MD2_METRICS_H_ */
