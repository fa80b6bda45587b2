"""On-device input pipeline: the colour pyramid the loss reads, built on the GPU from the
decoded uint8 frames instead of by PIL in 12 loader workers.

Mirrors KITTIMonoDataset_v2.__getitem__ / resize_intrinsic (model_loader/kitti_mono.py:283-288, 319-329,
347-362): flip, transforms.Resize(.., Image.ANTIALIAS) from the original image to every level,
transforms.ToTensor(), and the do_color branch (ColorJitter).  Bit-identical to Pillow / torchvision's PIL path
(include/md2_pipeline.h explains why).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import cabi

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = cabi.load_library()
    return _lib


def pyramid_tables(N, Hin, Win, H, W, scales=4):
    """Host-side Lanczos coefficient tables (int32 numpy array) - no GPU needed."""
    cfg = cabi.md2_pyramid_cfg(N, Hin, Win, H, W, scales)
    lib = _L()
    nbytes = lib.md2_pyramid_tables_bytes(C.byref(cfg))
    if nbytes == 0:
        raise RuntimeError(f"pyramid_tables: invalid configuration {(N, Hin, Win, H, W, scales)}")
    tab = np.empty(nbytes // 4, dtype=np.int32)
    rc = lib.md2_pyramid_tables_fill(C.byref(cfg), C.c_void_p(tab.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"md2_pyramid_tables_fill failed with code {rc}")
    return cfg, tab


class ColorPyramid:
    """pyr = ColorPyramid(N, 375, 1242, 192, 640, device="cuda:0"); levels = pyr(images_u8, flip)
    images_u8: uint8 CUDA tensor [N, Hin, Win, 3]; flip: None or bool/uint8 [N]; returns `scales` float32
    tensors [N, 3, H >> s, W >> s], i.e. inputs[("color", f, s)] for the N = batch x frames images."""

    def __init__(self, N, Hin, Win, height, width, scales=4, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ColorPyramid: CUDA device required; md2_b200 has no CPU path")
        self.cfg, tab = pyramid_tables(N, Hin, Win, height, width, scales)
        self.tables = torch.from_numpy(tab).to(self.device)
        self.workspace = torch.empty(_L().md2_pyramid_workspace_bytes(C.byref(self.cfg)), dtype=torch.uint8,
                                     device=self.device)
        self.shape = (N, Hin, Win, 3)

    def __call__(self, images, flip=None):
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8
                and tuple(images.shape) == self.shape):
            raise RuntimeError(f"ColorPyramid: expected a uint8 CUDA tensor {self.shape}, got "
                               f"{getattr(images, 'dtype', type(images))} {tuple(getattr(images, 'shape', ()))}")
        images = images.contiguous()
        if flip is not None:
            flip = flip.to(device=images.device, dtype=torch.uint8).contiguous()
            if flip.numel() != self.cfg.N:
                raise RuntimeError("ColorPyramid: flip must have one entry per image")
        c = self.cfg
        outs = [torch.empty(c.N, 3, c.H >> s, c.W >> s, dtype=torch.float32, device=images.device) for s in range(c.scales)]
        ptrs = (C.c_void_p * c.scales)(*[o.data_ptr() for o in outs])
        with torch.cuda.device(images.device):
            rc = _L().md2_color_pyramid(C.byref(c), C.c_void_p(images.data_ptr()),
                                        C.c_void_p(flip.data_ptr() if flip is not None else 0),
                                        C.c_void_p(self.tables.data_ptr()), ptrs, C.c_void_p(self.workspace.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"md2_color_pyramid failed with code {rc}")
        return outs


def to_tensor(images, out=None):
    """transforms.ToTensor() on the device (kitti_mono.py:283, applied at :352 / :364): uint8 CUDA tensors
    [N,H,W,3] (the loader's resized PIL images as bytes) -> float32 [N,3,H,W] = v / 255, bit-identical to
    torchvision's.  ``images`` is one tensor or a list (the levels of the target pyramid, the source frames, ...):
    the whole list is converted by ONE launch (chunks of 16 groups).  A training step that uploads bytes and calls
    this moves a quarter of the host-link traffic of uploading the loader's float tensors.  ``out``: optional
    preallocated float32 tensors of the matching shapes."""
    single = isinstance(images, torch.Tensor)
    imgs = [images] if single else list(images)
    if not imgs:
        return []
    dev = imgs[0].device
    for t in imgs:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 3
                and t.device == dev):
            raise RuntimeError("to_tensor: expected uint8 CUDA tensors [N,H,W,3] on one device; md2_b200 has no CPU path "
                               f"(got {getattr(t, 'dtype', type(t))} {tuple(getattr(t, 'shape', ()))})")
    imgs = [t.contiguous() for t in imgs]
    if out is None:
        outs = [torch.empty(t.shape[0], 3, t.shape[1], t.shape[2], dtype=torch.float32, device=dev) for t in imgs]
    else:
        outs = [out] if isinstance(out, torch.Tensor) else list(out)
        if len(outs) != len(imgs):
            raise RuntimeError("to_tensor: one output per input")
        for t, o in zip(imgs, outs):
            if not (o.is_cuda and o.device == dev and o.dtype == torch.float32 and o.is_contiguous()
                    and tuple(o.shape) == (t.shape[0], 3, t.shape[1], t.shape[2])):
                raise RuntimeError("to_tensor: out must be contiguous float32 [N,3,H,W] on the inputs' device")
    lib = _L()
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for k in range(0, len(imgs), cabi.MD2_TO_TENSOR_MAX):
            chunk = list(zip(imgs, outs))[k:k + cabi.MD2_TO_TENSOR_MAX]
            groups = (cabi.md2_u8_images * len(chunk))(*[
                cabi.md2_u8_images(t.data_ptr(), o.data_ptr(), t.shape[0], t.shape[1], t.shape[2]) for t, o in chunk])
            rc = lib.md2_to_tensor(len(chunk), groups, stream)
            if rc != 0:
                raise RuntimeError(f"md2_to_tensor failed with code {rc}")
    return outs[0] if single else outs


class ColorJitter:
    """The jitter of KITTIMonoDataset_v2 (kitti_mono.py:281-282): ``ColorJitter.get_params((0.8, 1.2), (0.8, 1.2),
    (0.8, 1.2), (-0.1, 0.1))`` draws the four factors and a shuffled order once, like torchvision <= 0.8 did, and
    the object is then called on pyramid levels: float32 CUDA [N,3,h,w] (v/255) -> the same, bit-identical to the PIL
    pipeline (uint8 between the four steps).  ``apply``: optional bool/uint8 [N], the per-sample do_color flag."""

    OPS = ("brightness", "contrast", "saturation", "hue")

    def __init__(self, order, brightness, contrast, saturation, hue):
        self.order = [int(k) for k in order] + [-1] * (4 - len(order))
        self.factors = (float(brightness), float(contrast), float(saturation), float(hue))

    @staticmethod
    def get_params(brightness, contrast, saturation, hue):
        import random
        order, f = [], [1.0, 1.0, 1.0, 0.0]
        for k, rng in enumerate((brightness, contrast, saturation, hue)):   # torchvision 0.8 ColorJitter.get_params
            if rng is not None:
                f[k] = random.uniform(rng[0], rng[1])
                order.append(k)
        random.shuffle(order)
        return ColorJitter(order, *f)

    def __call__(self, images, apply=None):
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.float32 and images.dim() == 4
                and images.shape[1] == 3):
            raise RuntimeError("ColorJitter: expected a float32 CUDA tensor [N,3,H,W]; md2_b200 has no CPU path")
        images = images.contiguous()
        N, _, H, W = images.shape
        cfg = cabi.md2_jitter_cfg(N, H, W, (C.c_int * 4)(*self.order), *self.factors)
        lib = _L()
        nbytes = lib.md2_jitter_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            raise RuntimeError(f"ColorJitter: invalid parameters {self.order} {self.factors}")
        if apply is not None:
            apply = apply.to(device=images.device, dtype=torch.uint8).contiguous()
            if apply.numel() != N:
                raise RuntimeError("ColorJitter: apply must have one entry per image")
        with torch.cuda.device(images.device):
            ws = torch.empty(nbytes, dtype=torch.uint8, device=images.device)
            out = torch.empty_like(images)
            rc = lib.md2_color_jitter(C.byref(cfg), C.c_void_p(images.data_ptr()),
                                      C.c_void_p(apply.data_ptr() if apply is not None else 0), C.c_void_p(out.data_ptr()),
                                      C.c_void_p(ws.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"md2_color_jitter failed with code {rc}")
        return out


def resize_intrinsic(width, height, scales=4, variant="row1_width"):
    """K / inv_K per level as the reference loaders build them (host, numpy): "row1_width" is
    KITTIMonoDataset_v2.resize_intrinsic (kitti_mono.py:319-329, both rows scaled by the WIDTH, floor division),
    "monodepth2" the KITTIMonoDataset variant (kitti_mono.py:203-206)."""
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    out = {}
    for s in range(scales):
        Kc = K.copy()
        if variant == "row1_width":
            Kc[0, :] = Kc[0, :] * width // (2 ** s)
            Kc[1, :] = Kc[1, :] * width // (2 ** s)
        elif variant == "monodepth2":
            Kc[0, :] *= width // (2 ** s)
            Kc[1, :] *= height // (2 ** s)
        else:
            raise ValueError(variant)
        out[("K", s)] = torch.from_numpy(Kc)
        out[("inv_K", s)] = torch.from_numpy(np.linalg.pinv(Kc))
    return out
