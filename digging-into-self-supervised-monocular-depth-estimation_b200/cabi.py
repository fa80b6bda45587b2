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
                ("seed", C.c_uint64),
                ("seed_dev", C.c_void_p)]


class md2_outputs(C.Structure):
    _fields_ = [("loss", C.c_void_p), ("per_pixel", C.c_void_p),
                ("argmin", C.c_void_p), ("depth", C.c_void_p)]


class md2_grads(C.Structure):
    _fields_ = [("grad_disp", C.c_void_p * MAX_SCALES),
                ("grad_T", C.c_void_p * MAX_SOURCES)]


# every symbol include/md2_loss.h declares
EXPORTS = ["md2_workspace_bytes", "md2_loss_forward", "md2_loss_forward_backward",
           "md2_loss_backward", "md2_pose_forward", "md2_pose_backward",
           "md2_launches_per_step", "md2_version", "md2_debug_warp", "md2_loss_forward_backward_timed",
           "md2_debug_div"]

# include/md2_ops.h: the symbol-level operators (name -> argument types; all return int)
_V, _I, _L, _D = C.c_void_p, C.c_int, C.c_longlong, C.c_double
OPS_PROTOTYPES = {
    "md2_disp2depth_forward": [_L, _V, _D, _D, _V, _V, _V],
    "md2_disp2depth_backward": [_L, _V, _D, _D, _V, _V, _V, _V],
    "md2_upsample_forward": [_I, _I, _I, _I, _I, _V, _V, _V],
    "md2_upsample_backward": [_I, _I, _I, _I, _I, _V, _V, _V],
    "md2_backproject_forward": [_I, _I, _I, _V, _V, _V, _V],
    "md2_backproject_backward": [_I, _I, _I, _V, _V, _V, _V],
    "md2_project_forward": [_I, _I, _I, _V, _V, _V, _D, _V, _V],
    "md2_project_backward": [_I, _I, _I, _V, _V, _V, _D, _V, _V, _V, _V],
    "md2_grid_sample_forward": [_I, _I, _I, _I, _I, _I, _V, _V, _V, _V],
    "md2_grid_sample_backward": [_I, _I, _I, _I, _I, _I, _V, _V, _V, _V, _V],
    "md2_reprojection_forward": [_I, _I, _I, _V, _V, _V, _V],
    "md2_reprojection_backward": [_I, _I, _I, _V, _V, _V, _V, _V],
    "md2_smooth_forward": [_I, _I, _I, _V, _V, _V, _V, _V],
    "md2_smooth_backward": [_I, _I, _I, _V, _V, _V, _V, _V, _V],
    "md2_mean_inv_depth_forward": [_I, _I, _V, _V, _V],
    "md2_mean_inv_depth_backward": [_I, _I, _V, _V, _V, _V],
    "md2_reflection_pad2d_forward": [_I] * 9 + [_V, _V, _V],
    "md2_reflection_pad2d_backward": [_I] * 9 + [_V, _V, _V],
    "md2_maxpool2d_nhwc_forward": [_I] * 7 + [_V, _V, _V, _V],
    "md2_maxpool2d_nhwc_backward": [_I] * 7 + [_V, _V, _V, _V],
}
EXPORTS += list(OPS_PROTOTYPES) + ["md2_metrics_workspace_bytes", "md2_depth_metrics", "md2_pyramid_tables_bytes",
                                   "md2_pyramid_tables_fill", "md2_pyramid_workspace_bytes", "md2_color_pyramid",
                                   "md2_jitter_workspace_bytes", "md2_color_jitter", "md2_to_tensor"]


class md2_jitter_cfg(C.Structure):
    """include/md2_pipeline.h"""
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("order", C.c_int * 4),
                ("brightness", C.c_double), ("contrast", C.c_double), ("saturation", C.c_double), ("hue", C.c_double)]


class md2_pyramid_cfg(C.Structure):
    """include/md2_pipeline.h"""
    _fields_ = [("N", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("H", C.c_int), ("W", C.c_int), ("scales", C.c_int)]


class md2_u8_images(C.Structure):
    """include/md2_pipeline.h"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int)]


MD2_TO_TENSOR_MAX = 16


class md2_metrics_cfg(C.Structure):
    """include/md2_metrics.h"""
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Hg", C.c_int), ("Wg", C.c_int),
                ("y0", C.c_int), ("y1", C.c_int), ("x0", C.c_int), ("x1", C.c_int),
                ("min_depth", C.c_float), ("max_depth", C.c_float)]

MD2_ERR_NULL, MD2_ERR_SHAPE, MD2_ERR_CONFIG, MD2_ERR_WORKSPACE, MD2_ERR_NO_DEVICE = -1, -2, -3, -4, -5

LIB_NAME = "libmd2loss.so"


def lib_path():
    # MD2_LIB selects an alternative build of the same library (tile-shape experiments, tools/variants.py)
    return os.environ.get("MD2_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


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
    lib.md2_loss_forward_backward_timed.restype = C.c_int
    lib.md2_loss_forward_backward_timed.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs),
                                                    C.POINTER(md2_outputs), C.POINTER(md2_grads), C.c_float,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.md2_debug_div.restype = C.c_int
    lib.md2_debug_div.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.md2_debug_warp.restype = C.c_int
    lib.md2_debug_warp.argtypes = [C.POINTER(md2_cfg), C.POINTER(md2_inputs), C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.md2_metrics_workspace_bytes.restype = C.c_size_t
    lib.md2_metrics_workspace_bytes.argtypes = [C.POINTER(md2_metrics_cfg)]
    lib.md2_depth_metrics.restype = C.c_int
    lib.md2_depth_metrics.argtypes = [C.POINTER(md2_metrics_cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    for name in ("md2_pyramid_tables_bytes", "md2_pyramid_workspace_bytes"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.POINTER(md2_pyramid_cfg)]
    lib.md2_jitter_workspace_bytes.restype = C.c_size_t
    lib.md2_jitter_workspace_bytes.argtypes = [C.POINTER(md2_jitter_cfg)]
    lib.md2_color_jitter.restype = C.c_int
    lib.md2_color_jitter.argtypes = [C.POINTER(md2_jitter_cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.md2_pyramid_tables_fill.restype = C.c_int
    lib.md2_pyramid_tables_fill.argtypes = [C.POINTER(md2_pyramid_cfg), C.c_void_p]
    lib.md2_color_pyramid.restype = C.c_int
    lib.md2_color_pyramid.argtypes = [C.POINTER(md2_pyramid_cfg), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    lib.md2_to_tensor.restype = C.c_int
    lib.md2_to_tensor.argtypes = [C.c_int, C.POINTER(md2_u8_images), C.c_void_p]
    for name, argtypes in OPS_PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
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


class CLoss:
    """Thin convenience wrapper used by the GPU parity tests: calls the C ABI of
    libmd2loss.so with raw device pointers of torch CUDA tensors on the current stream."""

    def __init__(self, path=None):
        self.lib = load_library(path)

    @staticmethod
    def _prep(args):
        import torch
        f = lambda t: t.detach().to(torch.float32).contiguous()
        a = dict(target=f(args["target"]), sources=[f(t) for t in args["sources"]],
                 disps=[f(t) for t in args["disps"]], color_pyr=[f(t) for t in args["color_pyr"]],
                 K=f(args["K"]), inv_K=f(args["inv_K"]), Ts=[f(t) for t in args["Ts"]],
                 noise=[f(t) for t in args["noise"]] if args.get("noise") is not None else None)
        B, _, H, W = a["target"].shape
        cfg = make_cfg(B, H, W, len(a["sources"]), len(a["disps"]), args.get("automask", True),
                       args.get("min_depth", 0.1), args.get("max_depth", 100.0),
                       args.get("disp_smoothness", 1e-3))
        inp = make_inputs(a["target"], a["sources"], a["disps"], a["color_pyr"], a["K"], a["inv_K"],
                          a["Ts"], a["noise"], args.get("seed", 0))
        return a, cfg, inp

    def _ws(self, cfg, dev):
        import torch
        n = self.lib.md2_workspace_bytes(C.byref(cfg))
        assert n > 0
        return torch.empty(n, dtype=torch.uint8, device=dev)

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _alloc_out(self, cfg, dev):
        import torch
        ns, B, H, W = cfg.num_scales, cfg.B, cfg.H, cfg.W
        return dict(loss=torch.zeros(1, device=dev), per_pixel=torch.zeros(ns, B, H, W, device=dev),
                    argmin=torch.zeros(ns, B, H, W, dtype=torch.uint8, device=dev),
                    depth=torch.zeros(ns, B, 1, H, W, device=dev))

    def forward(self, args):
        a, cfg, inp = self._prep(args)
        dev = a["target"].device
        o = self._alloc_out(cfg, dev)
        ws = self._ws(cfg, dev)
        out = make_outputs(o["loss"], o["per_pixel"], o["argmin"], o["depth"])
        rc = self.lib.md2_loss_forward(C.byref(cfg), C.byref(inp), C.byref(out), C.c_void_p(ws.data_ptr()),
                                       self._stream())
        if rc != 0:
            raise RuntimeError(f"md2_loss_forward returned {rc}")
        return o

    def forward_backward(self, args, grad_loss=1.0, events=None):
        """events = (start, stop) torch.cuda.Event pair (already recorded once): timed around the tile kernel."""
        import torch
        a, cfg, inp = self._prep(args)
        dev = a["target"].device
        o = self._alloc_out(cfg, dev)
        ws = self._ws(cfg, dev)
        gd = [torch.full_like(d, float("nan")) for d in a["disps"]]
        gT = [torch.full((cfg.B, 4, 4), float("nan"), device=dev) for _ in a["Ts"]]
        out = make_outputs(o["loss"], o["per_pixel"], o["argmin"], o["depth"])
        g = make_grads(gd, gT)
        if events is not None:
            rc = self.lib.md2_loss_forward_backward_timed(
                C.byref(cfg), C.byref(inp), C.byref(out), C.byref(g), C.c_float(grad_loss), C.c_void_p(ws.data_ptr()),
                self._stream(), C.c_void_p(events[0].cuda_event), C.c_void_p(events[1].cuda_event))
        else:
            rc = self.lib.md2_loss_forward_backward(C.byref(cfg), C.byref(inp), C.byref(out), C.byref(g),
                                                    C.c_float(grad_loss), C.c_void_p(ws.data_ptr()), self._stream())
        if rc != 0:
            raise RuntimeError(f"md2_loss_forward_backward returned {rc}")
        o["grad_disp"], o["grad_T"] = gd, gT
        return o

    def backward(self, args, argmin, grad_loss):
        import torch
        a, cfg, inp = self._prep(args)
        dev = a["target"].device
        ws = self._ws(cfg, dev)
        gd = [torch.full_like(d, float("nan")) for d in a["disps"]]
        gT = [torch.full((cfg.B, 4, 4), float("nan"), device=dev) for _ in a["Ts"]]
        g = make_grads(gd, gT)
        gl = torch.as_tensor([float(grad_loss)], dtype=torch.float32, device=dev)
        am = argmin.contiguous()
        rc = self.lib.md2_loss_backward(C.byref(cfg), C.byref(inp), C.c_void_p(am.data_ptr()),
                                        C.c_void_p(gl.data_ptr()), C.byref(g), C.c_void_p(ws.data_ptr()),
                                        self._stream())
        if rc != 0:
            raise RuntimeError(f"md2_loss_backward returned {rc}")
        return dict(grad_disp=gd, grad_T=gT)

    def debug_warp(self, args, scale, source):
        import torch
        a, cfg, inp = self._prep(args)
        dev = a["target"].device
        ws = self._ws(cfg, dev)
        coords = torch.zeros(cfg.B, 2, cfg.H, cfg.W, device=dev)
        warped = torch.zeros(cfg.B, 3, cfg.H, cfg.W, device=dev)
        rc = self.lib.md2_debug_warp(C.byref(cfg), C.byref(inp), scale, source, C.c_void_p(coords.data_ptr()),
                                     C.c_void_p(warped.data_ptr()), C.c_void_p(ws.data_ptr()), self._stream())
        if rc != 0:
            raise RuntimeError(f"md2_debug_warp returned {rc}")
        return coords, warped
