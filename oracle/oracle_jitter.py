"""TEST INFRASTRUCTURE ONLY - numpy restatement of the colour-jitter branch of the reference's loader.

KITTIMonoDataset_v2 (model_loader/kitti_mono.py:281-282, 351-357) draws ONE jitter with
transforms.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)) in __init__ and applies it to
every resized PIL image of a sample when do_color is set.  With the torchvision the reference was written for
(<= 0.8) get_params returns a Compose of the four adjust_* Lambdas in a shuffled order; each runs on the PIL
image, i.e. on uint8 data, through third-party code that is not part of /root/reference:
  torchvision/transforms/functional_pil.py  adjust_brightness / _contrast / _saturation / _hue
  Pillow ImageEnhance (-> Image.blend, libImaging/Blend.c), Image.convert("L") (Convert.c rgb2l),
  Image.convert("HSV") / back (Convert.c rgb2hsv_row, hsv2rgb)
This file restates those published algorithms; tests/test_oracle_jitter.py pins every piece bit-exactly against
the installed Pillow 12.2.0 / torchvision 0.26 (whose F.adjust_* still take PIL images), the two HSV conversions
exhaustively over all 2^24 colours, and against tests/golden/jitter.npz.
"""
import numpy as np

OPS = ("brightness", "contrast", "saturation", "hue")


def blend(im1, im2, alpha):
    """Image.blend(im1, im2, alpha) on uint8 arrays (libImaging/Blend.c): float32 product and sum, truncation;
    outside [0, 1] the result is clipped."""
    d = (im2.astype(np.int32) - im1.astype(np.int32)).astype(np.float32)
    t = (im1.astype(np.float64) + (np.float32(alpha) * d).astype(np.float64)).astype(np.float32)  # one fp32 rounding
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def luma(img):
    """Image.convert("L") (Convert.c rgb2l): ITU-R 601-2, 16-bit fixed point."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def adjust_brightness(img, factor):
    return blend(np.zeros_like(img), img, factor)


def adjust_contrast(img, factor):
    mean = int(luma(img).astype(np.float64).sum() / luma(img).size + 0.5)   # int(ImageStat.Stat(L).mean[0] + 0.5)
    return blend(np.full_like(img, mean), img, factor)


def adjust_saturation(img, factor):
    L = luma(img)
    return blend(np.stack([L, L, L], -1), img, factor)


def rgb_to_hsv(img):
    """Convert.c rgb2hsv_row: float32 ratios, hue through double, truncation to uint8."""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = cr / maxc.astype(np.float32)
        rc = (maxc - r).astype(np.float32) / cr
        gc = (maxc - g).astype(np.float32) / cr
        bc = (maxc - b).astype(np.float32) / cr
        h = np.where(r == maxc, (bc - gc).astype(np.float32),
                     np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(np.float32),
                              (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(np.float32)))
        h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
        uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def _c_round(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def hsv_to_rgb(hsv):
    """Convert.c hsv2rgb: sector and remainder through double, C round() of the three blends."""
    h = hsv[..., 0].astype(np.float64)
    s = hsv[..., 1]
    v = hsv[..., 2].astype(np.int64)
    hf = h * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i).astype(np.float32).astype(np.float64)
    fs = (s.astype(np.float64) / 255.0).astype(np.float32).astype(np.float64)
    vf = v.astype(np.float64)
    p = np.clip(_c_round(vf * (1.0 - fs)), 0, 255).astype(np.int64)
    q = np.clip(_c_round(vf * (1.0 - fs * f)), 0, 255).astype(np.int64)
    t = np.clip(_c_round(vf * (1.0 - fs * (1.0 - f))), 0, 255).astype(np.int64)
    k = i % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    gray = s == 0
    return np.stack([np.where(gray, v, r), np.where(gray, v, g), np.where(gray, v, b)], -1).astype(np.uint8)


def hue_shift(factor):
    """functional_pil.adjust_hue: the uint8 added (with wrap-around) to the H channel."""
    return int(np.int32(factor * 255).astype(np.uint8))


def adjust_hue(img, factor):
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = hsv[..., 0] + np.uint8(hue_shift(factor))   # uint8 wrap-around
    return hsv_to_rgb(hsv)


def color_jitter(img, order, brightness, contrast, saturation, hue):
    """Apply the four adjustments to a uint8 [H,W,3] image in `order` (a permutation of 0..3 indexing OPS)."""
    fn = (lambda a: adjust_brightness(a, brightness), lambda a: adjust_contrast(a, contrast),
          lambda a: adjust_saturation(a, saturation), lambda a: adjust_hue(a, hue))
    for k in order:
        img = fn[k](img)
    return img
