/* md2_ops.h - C ABI of the reference's symbol-level (unfused) operators, SURVEY.md 8b "L1".
 *
 * The training path is the fused kernel of md2_loss.h.  These entry points cover a caller that composes
 * the operators itself (model_tool/processor.py:139-187 does, and its posecnn branch :153-157 derives a
 * per-scale pose from the depth map in between).  Same conventions as md2_loss.h: device pointers to
 * contiguous fp32 NCHW, the caller's stream, return 0 / MD2_ERR_* (<0) / cudaError_t (>0), no torch types,
 * no CPU path.  Each forward reproduces the rounding sequence of the ATen operators the reference calls;
 * each backward is the analytic adjoint (gradients overwrite their output buffer).
 */
#ifndef MD2_OPS_H_
#define MD2_OPS_H_

#include "md2_loss.h"

#ifdef __cplusplus
extern "C" {
#endif

/* disparity2depth (model_layer/warp.py:29-39): scaled = 1/max_depth + (1/min_depth - 1/max_depth) * disp,
 * depth = 1 / scaled.  `scaled` or `depth` may be NULL.  Backward: g_scaled / g_depth may be NULL (= 0). */
int md2_disp2depth_forward(long long n, const float* disp, double min_depth, double max_depth,
                           float* scaled, float* depth, md2_stream_t stream);
int md2_disp2depth_backward(long long n, const float* disp, double min_depth, double max_depth,
                            const float* g_scaled, const float* g_depth, float* g_disp, md2_stream_t stream);

/* interpolate(tensor, H, W, "bilinear", align_corners=False) (model_layer/warp.py:18-20) on `planes`
 * = N*C planes of h x w -> H x W.  Backward is the adjoint (g_in [planes,h,w] is overwritten). */
int md2_upsample_forward(int planes, int h, int w, int H, int W, const float* in, float* out, md2_stream_t stream);
int md2_upsample_backward(int planes, int h, int w, int H, int W, const float* g_out, float* g_in,
                          md2_stream_t stream);

/* Depth2PointCloud.forward (model_layer/warp.py:236-246): depth [B,1,H,W], inv_K [B,4,4] ->
 * camera points [B,4,H*W] (fourth row ones).  Backward returns dL/d depth. */
int md2_backproject_forward(int B, int H, int W, const float* depth, const float* inv_K, float* cam,
                            md2_stream_t stream);
int md2_backproject_backward(int B, int H, int W, const float* inv_K, const float* g_cam, float* g_depth,
                             md2_stream_t stream);

/* PointCloud2Pixel.forward (model_layer/warp.py:259-269): camera points [B,4,H*W], K, T [B,4,4] ->
 * normalised sampling grid [B,H,W,2].  Backward returns dL/d cam [B,4,H*W] and dL/dT [B,4,4]. */
int md2_project_forward(int B, int H, int W, const float* cam, const float* K, const float* T, double eps,
                        float* grid, md2_stream_t stream);
int md2_project_backward(int B, int H, int W, const float* cam, const float* K, const float* T, double eps,
                         const float* g_grid, float* g_cam, float* g_T, md2_stream_t stream);

/* grid_sample(tensor, coords, "border", align_corners=True), bilinear (model_layer/warp.py:12-14):
 * img [B,C,H,W], grid [B,Ho,Wo,2] -> out [B,C,Ho,Wo].  Backward returns dL/d grid (the image is data on
 * the reference's path: processor.py:172-176 samples inputs[("color", f, 0)]). */
int md2_grid_sample_forward(int B, int C, int H, int W, int Ho, int Wo, const float* img, const float* grid,
                            float* out, md2_stream_t stream);
int md2_grid_sample_backward(int B, int C, int H, int W, int Ho, int Wo, const float* img, const float* grid,
                             const float* g_out, float* g_grid, md2_stream_t stream);

/* ReprojectionLoss.forward (model_loss/model_loss.py:92-103): 0.85 * mean_c SSIM-dissimilarity +
 * 0.15 * mean_c |target - pred|, 3x3 reflection-padded windows; pred, target [B,3,H,W] -> [B,1,H,W].
 * Backward returns dL/d pred (the target is data). */
int md2_reprojection_forward(int B, int H, int W, const float* pred, const float* target, float* out,
                             md2_stream_t stream);
int md2_reprojection_backward(int B, int H, int W, const float* pred, const float* target, const float* g_out,
                              float* g_pred, md2_stream_t stream);

/* SmoothLoss.forward (model_loss/model_loss.py:107-116): mean-normalised edge-aware smoothness of
 * disp [B,1,h,w] against color [B,3,h,w] -> loss[1].  `part` is 3*B floats of scratch that the backward
 * reads again; g_loss_dev is a device pointer to the upstream scalar. */
int md2_smooth_forward(int B, int h, int w, const float* disp, const float* color, float* loss, float* part,
                       md2_stream_t stream);
int md2_smooth_backward(int B, int h, int w, const float* disp, const float* color, const float* part,
                        const float* g_loss_dev, float* g_disp, md2_stream_t stream);

/* (1 / depth).mean(3, True).mean(2, True) of the posecnn branch (model_tool/processor.py:155):
 * depth [B, n = H*W] -> out [B]; the backward returns dL/d depth from g_out [B]. */
int md2_mean_inv_depth_forward(int B, int n, const float* depth, float* out, md2_stream_t stream);
int md2_mean_inv_depth_backward(int B, int n, const float* depth, const float* g_out, float* g_depth,
                                md2_stream_t stream);

/* nn.ReflectionPad2d((pad_l, pad_r, pad_t, pad_b)) of the decoder's Conv3x3 blocks (model_layer/depth_decoder.py:36-50,
 * model_layer/warp.py:175-190), in the memory format the tensor already has: channels_last != 0 means the buffers are
 * NHWC ([N,H,W,C] in memory, torch.channels_last), 0 means contiguous NCHW.  ATen's operator only knows NCHW, which costs
 * a channels-last network two layout copies per pad and per direction; this one keeps NHWC tensors NHWC (float4 moves
 * when C % 4 == 0).  in [N,C,H,W] -> out [N,C,H+pad_t+pad_b,W+pad_l+pad_r], bit-identical to ATen; every pad must be
 * smaller than its axis (MD2_ERR_SHAPE otherwise, like torch's error).  Backward: g_in is overwritten with the sum of
 * the <= 9 padded positions that read each element, added in a fixed order (ATen accumulates with atomics). */
int md2_reflection_pad2d_forward(int N, int C, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                                 int channels_last, const float* in, float* out, md2_stream_t stream);
int md2_reflection_pad2d_backward(int N, int C, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b,
                                  int channels_last, const float* g_out, float* g_in, md2_stream_t stream);

/* nn.MaxPool2d(kernel, stride, padding) (dilation 1, ceil_mode False) of the ResNet encoders
 * (model_layer/depth_encoder.py:29: 3, 2, 1) on channels-last buffers: in [N,H,W,C] -> out [N,Ho,Wo,C] with
 * Ho = (H + 2*padding - kernel) / stride + 1, bit-identical to ATen (ties go to the first element in row-major window
 * order, NaN wins).  `winner` [N,Ho,Wo,C] uint8 receives each maximum's offset inside its window (dy*kernel + dx) and is
 * what the backward reads: g_in [N,H,W,C] is overwritten with the sum of the output gradients whose winner the element
 * is - a gather in a fixed order (ATen: atomics into a zero-filled buffer, 8-byte indices).  kernel <= 15,
 * 2*padding <= kernel. */
int md2_maxpool2d_nhwc_forward(int N, int C, int H, int W, int kernel, int stride, int padding, const float* in,
                               float* out, uint8_t* winner, md2_stream_t stream);
int md2_maxpool2d_nhwc_backward(int N, int C, int H, int W, int kernel, int stride, int padding, const float* g_out,
                                const uint8_t* winner, float* g_in, md2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MD2_OPS_H_ */
