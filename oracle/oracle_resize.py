"""TEST INFRASTRUCTURE ONLY - numpy restatement of the colour-pyramid step of the reference's loader.

KITTIMonoDataset_v2.__getitem__ (model_loader/kitti_mono.py:283-288, 296-304, 347-362) resizes the decoded
RGB image with transforms.Resize(.., interpolation=Image.ANTIALIAS) to every pyramid level and converts it
with transforms.ToTensor().  The resampler itself lives in a third-party dependency that is not part of
/root/reference: Pillow (12.2.0 in this image; src/libImaging/Resample.c - precompute_coeffs,
normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc).  This file restates
that published algorithm; tests/test_oracle_resize.py pins it bit-exactly against Pillow itself and against
the committed vectors tests/golden/pyramid.npz that tests/golden/make_golden_pyramid.py produced by calling the
reference dataset's own resize / ToTensor objects.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _lanczos(x):
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a = x * math.pi
        return (math.sin(a) / a) * (math.sin(a / 3.0) / (a / 3.0))
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: (bounds [out,2], coeffs [out,ksize] int32, ksize)."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 3.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resize_antialias(img, H, W):
    """uint8 [Hin,Win,3] -> uint8 [H,W,3]: horizontal pass to uint8, then vertical pass (ImagingResample)."""
    Hin, Win, _ = img.shape
    bx, kx, _ = precompute_coeffs(Win, W)
    by, ky, _ = precompute_coeffs(Hin, H)
    a = img.astype(np.int64)
    tmp = np.zeros((Hin, W, 3), np.uint8)
    for x in range(W):
        x0, n = bx[x]
        acc = (1 << (PRECISION_BITS - 1)) + (a[:, x0:x0 + n, :] * kx[x, :n, None]).sum(1)
        tmp[:, x, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    t = tmp.astype(np.int64)
    out = np.zeros((H, W, 3), np.uint8)
    for y in range(H):
        y0, n = by[y]
        acc = (1 << (PRECISION_BITS - 1)) + (t[y0:y0 + n] * ky[y, :n, None, None]).sum(0)
        out[y] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def to_tensor(img):
    """transforms.ToTensor() of kitti_mono.py:283 (applied at :352, :364) on uint8 images: [..., H, W, 3] uint8 ->
    [..., 3, H, W] float32, v / 255 (torchvision's to_tensor: permute, .to(float32), .div(255) - an IEEE division)."""
    a = np.asarray(img)
    assert a.dtype == np.uint8 and a.shape[-1] == 3
    return np.ascontiguousarray(np.moveaxis(a, -1, -3)).astype(np.float32) / np.float32(255)


def color_pyramid(img, H, W, scales=4, flip=False):
    """kitti_mono.py:296-304 (flip), :283-288 + :352-355 (resize every level from the original), ToTensor:
    uint8 [Hin,Win,3] -> list of float32 [3, H>>s, W>>s]."""
    if flip:
        img = img[:, ::-1, :]
    out = []
    for s in range(scales):
        r = resize_antialias(np.ascontiguousarray(img), H >> s, W >> s)
        out.append(np.ascontiguousarray(r.transpose(2, 0, 1)).astype(np.float32) / np.float32(255))
    return out
