"""ctypes mirror of include/md2_loss.h.

Used by the tests to call the C ABI of ``libmd2loss.so`` directly (raw device pointers
taken from torch tensors) and to check that the library exports every declared symbol.
The training path goes through the C++ extension (``csrc/torch_ext.cpp``) instead; both
end in the same ``extern "C"`` entry points.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_SOURCES = 4
MAX_SCALES = 4

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


class md2_cfg(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("S", C.c_int),
                ("num_scales", C.c_int), ("automask", C.c_int),
                ("min_depth", C.c_double), ("max_depth", C.c_double),
                ("disp_smoothness", C.c_double), ("eps_proj", C.c_double)]


class md2_inputs(C.Structure):
    _fields_ = [("target", C.c_void_p),
                ("sources", C.c_void_p * MAX_SOURCES),
                ("disp", C.c_void_p * MAX_SCALES),
                ("color_pyr", C.c_void_p * MAX_SCALES),
                ("K", C.c_void_p), ("inv_K", C.c_void_p),
                ("T", C.c_void_p * MAX_SOURCES),
                ("noise", C.c_void_p * MAX_SCALES),
                ("seed", C.c_uint64)]


class md2_outputs(C.Structure):
    _fields_ = [("loss", C.c_void_p), ("per_pixel", C.c_void_p),
                ("argmin", C.c_void_p), ("depth", C.c_void_p)]


class md2_grads(C.Structure):
    _fields_ = [("grad_disp", C.c_void_p * MAX_SCALES),
                ("grad_T", C.c_void_p * MAX_SOURCES)]


# every symbol include/md2_loss.h declares
EXPORTS = ["md2_workspace_bytes", "md2_loss_forward", "md2_loss_forward_backward",
           "md2_loss_backward", "md2_pose_forward", "md2_pose_backward",
           "md2_launches_per_step", "md2_version", "md2_debug_warp"]

LIB_NAME = "libmd2loss.so"


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def load_library(path=None):
    """dlopen the C-ABI library and declare its prototypes.  Raises if it is missing:
    there is no fallback implementation."""
    path = path or lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(path)
    lib.md2_workspace_bytes.restype = C.c_size_t
    lib.md2_workspace_bytes.argtypes = [C.POINTER(md2_cfg)]
    lib.md2_loss_forward.restype = C.c_int
    lib.md2_loss_forward.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs), C.POINTER(md2_outputs),
                                     C.c_void_p, C.c_void_p]
    lib.md2_loss_forward_backward.restype = C.c_int
    lib.md2_loss_forward_backward.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs),
                                              C.POINTER(md2_outputs), C.POINTER(md2_grads), C.c_float,
                                              C.c_void_p, C.c_void_p]
    lib.md2_loss_backward.restype = C.c_int
    lib.md2_loss_backward.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs), C.c_void_p, C.c_void_p,
                                      C.POINTER(md2_grads), C.c_void_p, C.c_void_p]
    lib.md2_pose_forward.restype = C.c_int
    lib.md2_pose_forward.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.md2_pose_backward.restype = C.c_int
    lib.md2_pose_backward.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    lib.md2_launches_per_step.restype = C.c_int
    lib.md2_launches_per_step.argtypes = [C.POINTER(md2_cfg), C.c_int]
    lib.md2_version.restype = C.c_char_p
    lib.md2_version.argtypes = []
    lib.md2_debug_warp.restype = C.c_int
    lib.md2_debug_warp.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs), C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def make_cfg(B, H, W, S, num_scales=4, automask=True, min_depth=0.1, max_depth=100.0,
             disp_smoothness=1e-3, eps_proj=1e-7):
    return md2_cfg(B, H, W, S, num_scales, int(bool(automask)), min_depth, max_depth,
                   disp_smoothness, eps_proj)


def make_inputs(target, sources, disps, color_pyr, K, inv_K, Ts, noise=None, seed=0):
    """Pack contiguous fp32 tensors (any device) into an md2_inputs.  The caller keeps the
    tensors alive for the duration of the call."""
    tensors = [target, K, inv_K] + list(sources) + list(disps) + list(color_pyr) + list(Ts) + \
              (list(noise) if noise is not None else [])
    for t in tensors:
        assert t.is_contiguous() and str(t.dtype) == "torch.float32", "fp32 contiguous tensors required"
    s = md2_inputs()
    s.target = _ptr(target)
    for i, t in enumerate(sources):
        s.sources[i] = t.data_ptr()
    for i, t in enumerate(Ts):
        s.T[i] = t.data_ptr()
    for i, t in enumerate(disps):
        s.disp[i] = t.data_ptr()
    for i, t in enumerate(color_pyr):
        s.color_pyr[i] = t.data_ptr()
    if noise is not None:
        for i, t in enumerate(noise):
            s.noise[i] = t.data_ptr()
    s.K = _ptr(K)
    s.inv_K = _ptr(inv_K)
    s.seed = seed
    return s


def make_outputs(loss, per_pixel=None, argmin=None, depth=None):
    return md2_outputs(_ptr(loss), _ptr(per_pixel), _ptr(argmin), _ptr(depth))


def make_grads(grad_disp, grad_T):
    g = md2_grads()
    for i, t in enumerate(grad_disp):
        g.grad_disp[i] = t.data_ptr()
    for i, t in enumerate(grad_T):
        g.grad_T[i] = t.data_ptr() if t is not None else None
    return g
