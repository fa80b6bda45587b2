"""Training-time depth metrics: drop-in for compute_depth_metric of
model_loss/model_metric.py:70-106 as called from model_tool/logger.py:30-36.

One C-ABI call (include/md2_metrics.h, csrc/md2_metrics.cu): 9 launches, exact medians by radix select,
no host synchronisation.  The seven metrics stay on the device; the reference's logger moves each to the
host with its own .cpu() - with this drop-in that is one sync for all of them if the caller reads the
packed tensor (``depth_metrics(...)``) instead.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi

_lib = None
METRIC_NAMES = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def _L():
    global _lib
    if _lib is None:
        _lib = cabi.load_library()
    return _lib


def depth_metrics(depth, gt, crop=(153, 371, 44, 1197), min_depth=1e-3, max_depth=80.0):
    """depth [B,1,H,W] (the fused loss's outputs[("depth", 0, 0)]), gt [B,1,Hg,Wg] -> float32 CUDA tensor [8]:
    abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, number of masked pixels."""
    for t, name in ((depth, "depth"), (gt, "gt")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == 1):
            raise RuntimeError(f"depth_metrics: {name} must be a float32 CUDA tensor [B,1,H,W]; md2_b200 has no CPU path")
    if depth.shape[0] != gt.shape[0] or depth.device != gt.device:
        raise RuntimeError("depth_metrics: depth and gt must share batch size and device")
    depth, gt = depth.detach().contiguous(), gt.detach().contiguous()
    B, _, H, W = depth.shape
    cfg = cabi.md2_metrics_cfg(B, H, W, gt.shape[2], gt.shape[3], crop[0], crop[1], crop[2], crop[3], min_depth, max_depth)
    lib = _L()
    nbytes = lib.md2_metrics_workspace_bytes(C.byref(cfg))
    if nbytes == 0:
        raise RuntimeError(f"depth_metrics: invalid configuration (crop {crop} outside a {gt.shape[2]}x{gt.shape[3]} ground truth?)")
    with torch.cuda.device(depth.device):
        ws = torch.empty(nbytes, dtype=torch.uint8, device=depth.device)
        out = torch.empty(8, dtype=torch.float32, device=depth.device)
        rc = lib.md2_depth_metrics(C.byref(cfg), C.c_void_p(depth.data_ptr()), C.c_void_p(gt.data_ptr()),
                                   C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"md2_depth_metrics failed with code {rc}")
    return out


def compute_depth_metric(inputs, outputs, lib="torch"):
    """model_metric.py:70-106: reads outputs[("depth", 0, 0)] and inputs[("depth", 0)], returns the tuple
    (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3) of 0-dim CUDA tensors."""
    if lib != "torch":
        raise NotImplementedError("md2_b200.compute_depth_metric: only lib='torch' (logger.py:33)")
    return tuple(depth_metrics(outputs[("depth", 0, 0)], inputs[("depth", 0)])[:7].unbind(0))
