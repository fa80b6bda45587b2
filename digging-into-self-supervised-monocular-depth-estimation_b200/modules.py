"""Symbol-level drop-ins (SURVEY.md 8b "L1"): the reference's unfused operators, same names and
signatures, each backed by a kernel of csrc/md2_l1.cu through the C ABI of include/md2_ops.h.

  model_layer/warp.py        grid_sample :12, interpolate :18, disparity2depth :29, param2matrix :126,
                             Depth2PointCloud :193, PointCloud2Pixel :250
  model_loss/model_loss.py   ReprojectionLoss :92, SmoothLoss :107

The training path does not go through these (it uses functional.view_synthesis_loss, one fused kernel);
they serve a caller that composes the operators itself.  Every forward reproduces the rounding sequence of
the ATen operators the reference calls, every backward is an analytic kernel registered with autograd.
fp32 CUDA tensors only; anything else raises (there is no PyTorch or CPU implementation behind them).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import cabi
from .functional import param2matrix  # noqa: F401  (warp.py:126-153)

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = cabi.load_library()
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError(f"{name}: expected a float32 CUDA tensor (got {getattr(t, 'dtype', type(t))} on "
                           f"{getattr(t, 'device', '?')}); md2_b200 has no CPU or mixed-precision path")
    return t.contiguous()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}")


# ----------------------------------------------------------------------------- disparity2depth
class _Disp2Depth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        disp = _f32(disp, "disparity2depth")
        scaled, depth = torch.empty_like(disp), torch.empty_like(disp)
        with torch.cuda.device(disp.device):
            _check(_L().md2_disp2depth_forward(disp.numel(), _p(disp), min_depth, max_depth, _p(scaled), _p(depth),
                                               _st()), "md2_disp2depth_forward")
        ctx.save_for_backward(disp)
        ctx.rng = (min_depth, max_depth)
        return scaled, depth

    @staticmethod
    def backward(ctx, g_scaled, g_depth):
        (disp,) = ctx.saved_tensors
        g = torch.empty_like(disp)
        gs = g_scaled.contiguous() if g_scaled is not None else None
        gd = g_depth.contiguous() if g_depth is not None else None
        with torch.cuda.device(disp.device):
            _check(_L().md2_disp2depth_backward(disp.numel(), _p(disp), ctx.rng[0], ctx.rng[1], _p(gs), _p(gd), _p(g),
                                                _st()), "md2_disp2depth_backward")
        return g, None, None


def disparity2depth(disparity, min_depth, max_depth):
    """warp.py:29-39 -> (scaled_disp, depth)."""
    return _Disp2Depth.apply(disparity, float(min_depth), float(max_depth))


# ----------------------------------------------------------------------------- interpolate
class _Upsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        x = _f32(x, "interpolate")
        N, Cc, h, w = x.shape
        out = torch.empty(N, Cc, H, W, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _check(_L().md2_upsample_forward(N * Cc, h, w, H, W, _p(x), _p(out), _st()), "md2_upsample_forward")
        ctx.shape = (N, Cc, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        N, Cc, h, w, H, W = ctx.shape
        g = g.contiguous()
        gi = torch.empty(N, Cc, h, w, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _check(_L().md2_upsample_backward(N * Cc, h, w, H, W, _p(g), _p(gi), _st()), "md2_upsample_backward")
        return gi, None, None


def interpolate(tensor, height, width, mode, align_corners):
    """warp.py:18-20; the reference only ever asks for ("bilinear", False) (processor.py:146)."""
    if mode != "bilinear" or align_corners:
        raise NotImplementedError("md2_b200.interpolate: only mode='bilinear', align_corners=False (processor.py:146)")
    return _Upsample.apply(tensor, int(height), int(width))


# ----------------------------------------------------------------------------- grid_sample
class _GridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, grid):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("md2_b200.grid_sample: the sampled image is data on the reference's path "
                                      "(processor.py:172-176); no gradient with respect to it")
        img, grid = _f32(img, "grid_sample"), _f32(grid, "grid_sample")
        B, Cc, H, W = img.shape
        if grid.dim() != 4 or grid.shape[0] != B or grid.shape[3] != 2:
            raise RuntimeError(f"grid_sample: coords must be [B,Ho,Wo,2], got {tuple(grid.shape)}")
        Ho, Wo = grid.shape[1], grid.shape[2]
        out = torch.empty(B, Cc, Ho, Wo, device=img.device, dtype=torch.float32)
        with torch.cuda.device(img.device):
            _check(_L().md2_grid_sample_forward(B, Cc, H, W, Ho, Wo, _p(img), _p(grid), _p(out), _st()),
                   "md2_grid_sample_forward")
        ctx.save_for_backward(img, grid)
        return out

    @staticmethod
    def backward(ctx, g):
        img, grid = ctx.saved_tensors
        B, Cc, H, W = img.shape
        g = g.contiguous()
        gg = torch.empty_like(grid)
        with torch.cuda.device(img.device):
            _check(_L().md2_grid_sample_backward(B, Cc, H, W, grid.shape[1], grid.shape[2], _p(img), _p(grid), _p(g),
                                                 _p(gg), _st()), "md2_grid_sample_backward")
        return None, gg


def grid_sample(tensor, coords, padding_mode, align_corners):
    """warp.py:12-14; the reference only ever asks for ("border", True) (processor.py:172-176)."""
    if padding_mode != "border" or not align_corners:
        raise NotImplementedError("md2_b200.grid_sample: only padding_mode='border', align_corners=True")
    return _GridSample.apply(tensor, coords)


# ----------------------------------------------------------------------------- Depth2PointCloud
class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, B, H, W):
        depth, inv_K = _f32(depth, "Depth2PointCloud"), _f32(inv_K, "Depth2PointCloud")
        if depth.numel() != B * H * W or tuple(inv_K.shape) != (B, 4, 4):
            raise RuntimeError(f"Depth2PointCloud: depth {tuple(depth.shape)} / inv_K {tuple(inv_K.shape)} do not "
                               f"match batch {B}, {H}x{W}")
        cam = torch.empty(B, 4, H * W, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            _check(_L().md2_backproject_forward(B, H, W, _p(depth), _p(inv_K), _p(cam), _st()), "md2_backproject_forward")
        ctx.save_for_backward(inv_K)
        ctx.dims = (B, H, W, depth.shape)
        return cam

    @staticmethod
    def backward(ctx, g):
        (inv_K,) = ctx.saved_tensors
        B, H, W, shape = ctx.dims
        g = g.contiguous()
        gd = torch.empty(shape, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _check(_L().md2_backproject_backward(B, H, W, _p(inv_K), _p(g), _p(gd), _st()), "md2_backproject_backward")
        return gd, None, None, None, None


class Depth2PointCloud(nn.Module):
    """warp.py:193-246: depth [B,1,H,W], inv_K [B,4,4] -> homogeneous camera points [B,4,H*W].  The pixel grid
    the reference keeps as a parameter is generated inside the kernel."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inverse_intrinsic_matrix):
        return _Backproject.apply(depth, inverse_intrinsic_matrix, self.batch_size, self.height, self.width)


# ----------------------------------------------------------------------------- PointCloud2Pixel
class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam, K, T, B, H, W, eps):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("md2_b200.PointCloud2Pixel: the intrinsics are data (processor.py:166-170)")
        cam, K, T = _f32(cam, "PointCloud2Pixel"), _f32(K, "PointCloud2Pixel"), _f32(T, "PointCloud2Pixel")
        if tuple(cam.shape) != (B, 4, H * W) or tuple(K.shape) != (B, 4, 4) or tuple(T.shape) != (B, 4, 4):
            raise RuntimeError(f"PointCloud2Pixel: cam {tuple(cam.shape)}, K {tuple(K.shape)}, T {tuple(T.shape)} do "
                               f"not match batch {B}, {H}x{W}")
        grid = torch.empty(B, H, W, 2, device=cam.device, dtype=torch.float32)
        with torch.cuda.device(cam.device):
            _check(_L().md2_project_forward(B, H, W, _p(cam), _p(K), _p(T), eps, _p(grid), _st()), "md2_project_forward")
        ctx.save_for_backward(cam, K, T)
        ctx.dims = (B, H, W, eps)
        return grid

    @staticmethod
    def backward(ctx, g):
        cam, K, T = ctx.saved_tensors
        B, H, W, eps = ctx.dims
        g = g.contiguous()
        g_cam, g_T = torch.empty_like(cam), torch.empty_like(T)
        with torch.cuda.device(cam.device):
            _check(_L().md2_project_backward(B, H, W, _p(cam), _p(K), _p(T), eps, _p(g), _p(g_cam), _p(g_T), _st()),
                   "md2_project_backward")
        return g_cam, None, g_T, None, None, None, None


class PointCloud2Pixel(nn.Module):
    """warp.py:250-269: camera points [B,4,H*W], K, T [B,4,4] -> normalised sampling grid [B,H,W,2]."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, camera_coords, intrinsic_matrix, transformation_matrix):
        return _Project.apply(camera_coords, intrinsic_matrix, transformation_matrix, self.batch_size, self.height,
                              self.width, float(self.eps))


# ----------------------------------------------------------------------------- ReprojectionLoss
class _Reprojection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("md2_b200.ReprojectionLoss: the target is data (processor.py:192-204)")
        pred, target = _f32(pred, "ReprojectionLoss"), _f32(target, "ReprojectionLoss")
        if pred.shape != target.shape or pred.dim() != 4 or pred.shape[1] != 3:
            raise RuntimeError(f"ReprojectionLoss: expected two [B,3,H,W] tensors, got {tuple(pred.shape)} and "
                               f"{tuple(target.shape)}")
        B, _, H, W = pred.shape
        out = torch.empty(B, 1, H, W, device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            _check(_L().md2_reprojection_forward(B, H, W, _p(pred), _p(target), _p(out), _st()), "md2_reprojection_forward")
        ctx.save_for_backward(pred, target)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        B, _, H, W = pred.shape
        g = g.contiguous()
        gp = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            _check(_L().md2_reprojection_backward(B, H, W, _p(pred), _p(target), _p(g), _p(gp), _st()),
                   "md2_reprojection_backward")
        return gp, None


class ReprojectionLoss(nn.Module):
    """model_loss.py:92-103: 0.85 * SSIM dissimilarity + 0.15 * L1, channel means -> [B,1,H,W]."""

    def forward(self, prediction, target):
        return _Reprojection.apply(prediction, target)


# ----------------------------------------------------------------------------- SmoothLoss
class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, color):
        disp, color = _f32(disp, "SmoothLoss"), _f32(color, "SmoothLoss")
        B, c1, h, w = disp.shape
        if c1 != 1 or tuple(color.shape) != (B, 3, h, w):
            raise RuntimeError(f"SmoothLoss: expected disp [B,1,h,w] and color [B,3,h,w], got {tuple(disp.shape)} and "
                               f"{tuple(color.shape)}")
        loss = torch.empty(1, device=disp.device, dtype=torch.float32)
        part = torch.empty(B * 3, device=disp.device, dtype=torch.float32)
        with torch.cuda.device(disp.device):
            _check(_L().md2_smooth_forward(B, h, w, _p(disp), _p(color), _p(loss), _p(part), _st()), "md2_smooth_forward")
        ctx.save_for_backward(disp, color, part)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        disp, color, part = ctx.saved_tensors
        B, _, h, w = disp.shape
        g = g.reshape(1).contiguous()
        gd = torch.empty_like(disp)
        with torch.cuda.device(disp.device):
            _check(_L().md2_smooth_backward(B, h, w, _p(disp), _p(color), _p(part), _p(g), _p(gd), _st()),
                   "md2_smooth_backward")
        return gd, None


class SmoothLoss(nn.Module):
    """model_loss.py:107-116: mean-normalised disparity, edge-aware first differences -> 0-dim loss."""

    def forward(self, disp, color):
        return _Smooth.apply(disp, color)


# ----------------------------------------------------------------------------- mean inverse depth (posecnn)
class _MeanInvDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth):
        depth = _f32(depth, "mean_inv_depth")
        B = depth.shape[0]
        out = torch.empty(B, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            _check(_L().md2_mean_inv_depth_forward(B, depth.numel() // B, _p(depth), _p(out), _st()),
                   "md2_mean_inv_depth_forward")
        ctx.save_for_backward(depth)
        return out.view(B, 1, 1, 1)

    @staticmethod
    def backward(ctx, g):
        (depth,) = ctx.saved_tensors
        B = depth.shape[0]
        g = g.reshape(B).contiguous()
        gd = torch.empty_like(depth)
        with torch.cuda.device(depth.device):
            _check(_L().md2_mean_inv_depth_backward(B, depth.numel() // B, _p(depth), _p(g), _p(gd), _st()),
                   "md2_mean_inv_depth_backward")
        return gd


def mean_inv_depth(depth):
    """(1 / depth).mean(3, True).mean(2, True) of processor.py:155 -> [B,1,1,1]."""
    return _MeanInvDepth.apply(depth)


# ----------------------------------------------------------------------------- ReflectionPad2d (decoder Conv3x3)
class _ReflectionPad2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise RuntimeError(f"ReflectionPad2d: expected a 4-D float32 CUDA tensor (got {getattr(x, 'dtype', type(x))} "
                               f"{tuple(getattr(x, 'shape', ()))} on {getattr(x, 'device', '?')}); md2_b200 has no CPU path")
        N, Cc, H, W = x.shape
        pl, pr, pt, pb = pad
        if min(pad) < 0 or pl >= W or pr >= W or pt >= H or pb >= H:
            raise RuntimeError(f"ReflectionPad2d: padding {pad} must be non-negative and smaller than the input "
                               f"dimensions {(H, W)}")
        # NHWC stays NHWC, anything else is handled as contiguous NCHW
        cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        x = x if cl else x.contiguous()
        out = torch.empty((N, Cc, H + pt + pb, W + pl + pr), device=x.device, dtype=torch.float32,
                          memory_format=torch.channels_last if cl else torch.contiguous_format)
        if out.numel():
            with torch.cuda.device(x.device):
                _check(_L().md2_reflection_pad2d_forward(N, Cc, H, W, pl, pr, pt, pb, int(cl), _p(x), _p(out), _st()),
                       "md2_reflection_pad2d_forward")
        ctx.cfg = (N, Cc, H, W, pad, cl)
        return out

    @staticmethod
    def backward(ctx, g):
        N, Cc, H, W, (pl, pr, pt, pb), cl = ctx.cfg
        fmt = torch.channels_last if cl else torch.contiguous_format
        g = g.contiguous(memory_format=fmt)
        gi = torch.empty((N, Cc, H, W), device=g.device, dtype=torch.float32, memory_format=fmt)
        if gi.numel():
            with torch.cuda.device(g.device):
                _check(_L().md2_reflection_pad2d_backward(N, Cc, H, W, pl, pr, pt, pb, int(cl), _p(g), _p(gi), _st()),
                       "md2_reflection_pad2d_backward")
        return gi, None


class ReflectionPad2d(nn.Module):
    """nn.ReflectionPad2d for the decoder's Conv3x3 (depth_decoder.py:40, warp.py:179) that keeps the tensor's memory
    format: a channels-last input gives a channels-last output (ATen's operator converts to NCHW and back, forward
    and backward).  ``padding``: int or (left, right, top, bottom).  Drop-in:
    ``conv3x3.pad = md2_b200.modules.ReflectionPad2d(1)``; values bit-identical to nn.ReflectionPad2d."""

    def __init__(self, padding):
        super().__init__()
        self.padding = (int(padding),) * 4 if isinstance(padding, int) else tuple(int(p) for p in padding)
        if len(self.padding) != 4:
            raise ValueError("ReflectionPad2d: padding must be an int or (left, right, top, bottom)")

    def forward(self, x):
        return _ReflectionPad2d.apply(x, self.padding)

    def extra_repr(self):
        return f"{self.padding}"


# ----------------------------------------------------------------------------- MaxPool2d (encoder stem)
class _MaxPool2dNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        N, Cc, H, W = x.shape
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        if H + 2 * p < k or W + 2 * p < k or Ho <= 0 or Wo <= 0:
            raise RuntimeError(f"MaxPool2d: input {(H, W)} too small for kernel {k}, padding {p}")
        out = torch.empty((N, Cc, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        win = torch.empty((N, Cc, Ho, Wo), device=x.device, dtype=torch.uint8, memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            _check(_L().md2_maxpool2d_nhwc_forward(N, Cc, H, W, k, s, p, _p(x), _p(out), _p(win), _st()),
                   "md2_maxpool2d_nhwc_forward")
        ctx.save_for_backward(win)
        ctx.cfg = (N, Cc, H, W, k, s, p)
        return out

    @staticmethod
    def backward(ctx, g):
        (win,) = ctx.saved_tensors
        N, Cc, H, W, k, s, p = ctx.cfg
        g = g.contiguous(memory_format=torch.channels_last)
        gi = torch.empty((N, Cc, H, W), device=g.device, dtype=torch.float32, memory_format=torch.channels_last)
        with torch.cuda.device(g.device):
            _check(_L().md2_maxpool2d_nhwc_backward(N, Cc, H, W, k, s, p, _p(g), _p(win), _p(gi), _st()),
                   "md2_maxpool2d_nhwc_backward")
        return gi, None, None, None


class MaxPool2d(nn.Module):
    """nn.MaxPool2d(kernel_size, stride, padding) of the ResNet encoders (depth_encoder.py:29) for channels-last
    tensors: bit-identical values, one byte of state per output instead of an int64 index, and a backward that gathers
    instead of scattering with atomics.  Square kernel / stride / padding, dilation 1, ceil_mode False.  A tensor that
    is not channels-last (or has C == 1, where the formats coincide) is converted first: this module is meant for
    networks run in torch.channels_last."""

    def __init__(self, kernel_size, stride=None, padding=0):
        super().__init__()
        self.kernel_size, self.padding = int(kernel_size), int(padding)
        self.stride = int(stride) if stride is not None else int(kernel_size)
        if not (0 < self.kernel_size <= 15 and self.stride > 0 and 0 <= 2 * self.padding <= self.kernel_size):
            raise ValueError("MaxPool2d: kernel_size in 1..15, stride > 0, padding <= kernel_size / 2")

    def forward(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise RuntimeError(f"MaxPool2d: expected a 4-D float32 CUDA tensor (got {getattr(x, 'dtype', type(x))} "
                               f"{tuple(getattr(x, 'shape', ()))} on {getattr(x, 'device', '?')}); md2_b200 has no CPU path")
        return _MaxPool2dNHWC.apply(x.contiguous(memory_format=torch.channels_last), self.kernel_size, self.stride,
                                    self.padding)

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}"


def _square(v):
    if isinstance(v, int):
        return v
    v = tuple(v)
    return v[0] if len(set(v)) == 1 else None


def use_channels_last_pooling(module: nn.Module) -> int:
    """Replace every plain nn.MaxPool2d inside ``module`` (square window, dilation 1, ceil_mode False, no indices
    returned - the encoders' stem, depth_encoder.py:29) by MaxPool2d above; returns how many were replaced."""
    n = 0
    for parent in module.modules():
        for name, child in list(parent.named_children()):
            if type(child) is not nn.MaxPool2d or child.ceil_mode or child.return_indices:
                continue
            k, s, p, d = (_square(child.kernel_size), _square(child.stride if child.stride is not None else child.kernel_size),
                          _square(child.padding), _square(child.dilation))
            if None in (k, s, p) or d != 1 or k > 15 or 2 * p > k:
                continue
            setattr(parent, name, MaxPool2d(k, s, p))
            n += 1
    return n


def use_channels_last_padding(module: nn.Module) -> int:
    """Replace every nn.ReflectionPad2d inside ``module`` (the reference's Conv3x3 blocks) by ReflectionPad2d above;
    returns how many were replaced.  Call it next to ``module.to(memory_format=torch.channels_last)``."""
    n = 0
    for parent in module.modules():
        for name, child in list(parent.named_children()):
            if type(child) is nn.ReflectionPad2d:
                setattr(parent, name, ReflectionPad2d(tuple(child.padding)))
                n += 1
    return n
