/* md2_pipeline.h - C ABI of the on-device colour pyramid, SURVEY.md 8f row N4 (the step before the loss).
 *
 * Replaces, per sample and frame, the CPU work of KITTIMonoDataset_v2.__getitem__
 * (model_loader/kitti_mono.py:283-288, 296-304, 347-362): optional left-right flip of the decoded RGB
 * image, transforms.Resize(.., interpolation=Image.ANTIALIAS) from the ORIGINAL image to each of the
 * `scales` pyramid levels (H >> s, W >> s), and transforms.ToTensor() (uint8 HWC -> float32 CHW / 255).
 * The colour-jitter branch (do_color) is md2_color_jitter at the end of this header.
 *
 * The resampler is Pillow's (third-party, not part of /root/reference; Pillow 12.2.0,
 * src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
 * Vertical_8bpc): separable Lanczos-3 whose support widens with the down-scaling factor, coefficients
 * normalised in double precision and rounded to 22-bit fixed point, a horizontal pass to uint8 and then
 * a vertical pass, each with +0.5 rounding and clipping.  It is integer arithmetic, so the results are
 * bit-identical to Pillow's.  The double-precision coefficient tables are computed on the HOST by
 * md2_pyramid_tables_fill (same libm as Pillow), uploaded once by the caller and reused for every batch.
 *
 * Conventions as in md2_loss.h.  images: uint8 [N, Hin, Win, 3] (decoded RGB, device); flip: uint8 [N]
 * or NULL; out[s]: float32 [N, 3, H >> s, W >> s].
 */
#ifndef MD2_PIPELINE_H_
#define MD2_PIPELINE_H_

#include "md2_loss.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct md2_pyramid_cfg {
  int N;        /* images in the call (batch x frames)           */
  int Hin, Win; /* decoded image size (375 x 1242 for KITTI)      */
  int H, W;     /* level-0 size (192 x 640)                       */
  int scales;   /* pyramid levels, 1..MD2_MAX_SCALES              */
} md2_pyramid_cfg;

/* Bytes of the coefficient tables (host and device copies have the same layout); 0 on an invalid cfg. */
size_t md2_pyramid_tables_bytes(const md2_pyramid_cfg* cfg);
/* Fill `host_tables` (md2_pyramid_tables_bytes bytes of HOST memory).  No GPU needed. */
int md2_pyramid_tables_fill(const md2_pyramid_cfg* cfg, void* host_tables);
/* Bytes of device scratch (the uint8 image between the two passes). */
size_t md2_pyramid_workspace_bytes(const md2_pyramid_cfg* cfg);

int md2_color_pyramid(const md2_pyramid_cfg* cfg, const uint8_t* images, const uint8_t* flip,
                      const void* device_tables, float* const* out /* [scales] host array of device pointers */,
                      void* workspace, md2_stream_t stream);

/* Colour jitter of the do_color branch (model_loader/kitti_mono.py:281-282, 351-357): the four adjustments of
 * torchvision's ColorJitter on uint8 images - brightness, contrast, saturation (Pillow ImageEnhance = Image.blend
 * with a degenerate image), hue (Pillow RGB -> HSV, uint8 wrap-around shift, HSV -> RGB) - applied in `order`
 * (op ids 0 brightness, 1 contrast, 2 saturation, 3 hue; -1 = skip), uint8 between the steps exactly like the PIL
 * pipeline, bit-identical to it.  in / out: float32 [N,3,H,W] holding v/255 (a pyramid level); apply: uint8 [N]
 * (0 = copy the image unchanged, the reference's do_color flag) or NULL. */
typedef struct md2_jitter_cfg {
  int N, H, W;
  int order[4];
  double brightness, contrast, saturation, hue; /* the Python floats torchvision draws; hue in [-0.5, 0.5] */
} md2_jitter_cfg;

size_t md2_jitter_workspace_bytes(const md2_jitter_cfg* cfg);
int md2_color_jitter(const md2_jitter_cfg* cfg, const float* in, const uint8_t* apply, float* out, void* workspace,
                     md2_stream_t stream);

/* transforms.ToTensor() on the device (model_loader/kitti_mono.py:283, applied at :352 and :364; kitti_stereo.py:50,
 * :130-153): uint8 [N,H,W,3] (the resized PIL image as bytes) -> float32 [N,3,H,W] = v / 255 correctly rounded, i.e.
 * bit-identical to torchvision's to_tensor (third-party; `img.permute(2,0,1).to(float32).div(255)`).  The loader keeps
 * bytes, the training step uploads a quarter of the float traffic and converts here.  One launch for up to
 * MD2_TO_TENSOR_MAX image groups (the target pyramid's levels, the source frames, ...); `groups` is a HOST array.
 * Groups with N*H*W == 0 are skipped. */
#define MD2_TO_TENSOR_MAX 16
typedef struct md2_u8_images {
  const uint8_t* src; /* device, [N,H,W,3] uint8  */
  float* dst;         /* device, [N,3,H,W] float32 */
  int N, H, W;
} md2_u8_images;
int md2_to_tensor(int count, const md2_u8_images* groups, md2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MD2_PIPELINE_H_ */
