"""Python front-end of the host emulation (tests only; see md2_emu.cpp)."""
import ctypes as C
import os
import subprocess

import torch

import md2_b200.cabi as cabi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "digging-into-self-supervised-monocular-depth-estimation_b200", "csrc")
LIB = os.path.join(HERE, "libmd2emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "md2_emu.cpp")] + [os.path.join(CSRC, f) for f in
                                                   ("md2_tile.cuh", "md2_platform.h", "md2_host.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", CSRC,
           srcs[0], "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for n in ("md2_emu_forward", "md2_emu_forward_backward", "md2_emu_backward", "md2_emu_debug_warp",
                  "md2_emu_pose_forward", "md2_emu_pose_backward"):
            getattr(_lib, n).restype = C.c_int
    return _lib


def _prep(args):
    f = lambda t: t.detach().to(torch.float32).contiguous()
    a = dict(target=f(args["target"]), sources=[f(t) for t in args["sources"]],
             disps=[f(t) for t in args["disps"]], color_pyr=[f(t) for t in args["color_pyr"]],
             K=f(args["K"]), inv_K=f(args["inv_K"]), Ts=[f(t) for t in args["Ts"]],
             noise=[f(t) for t in args["noise"]] if args.get("noise") is not None else None)
    B, _, H, W = a["target"].shape
    cfg = cabi.make_cfg(B, H, W, len(a["sources"]), len(a["disps"]), args.get("automask", True),
                        args.get("min_depth", 0.1), args.get("max_depth", 100.0),
                        args.get("disp_smoothness", 1e-3))
    inp = cabi.make_inputs(a["target"], a["sources"], a["disps"], a["color_pyr"], a["K"], a["inv_K"],
                           a["Ts"], a["noise"], args.get("seed", 0))
    return a, cfg, inp


def _alloc_out(cfg):
    ns, B, H, W = cfg.num_scales, cfg.B, cfg.H, cfg.W
    return dict(loss=torch.zeros(1), per_pixel=torch.zeros(ns, B, H, W),
                argmin=torch.zeros(ns, B, H, W, dtype=torch.uint8), depth=torch.zeros(ns, B, 1, H, W))


def _alloc_grads(a):
    return ([torch.full_like(d, float("nan")) for d in a["disps"]],
            [torch.full((T.shape[0], 4, 4), float("nan")) for T in a["Ts"]])


def forward(args):
    a, cfg, inp = _prep(args)
    o = _alloc_out(cfg)
    out = cabi.make_outputs(o["loss"], o["per_pixel"], o["argmin"], o["depth"])
    rc = lib().md2_emu_forward(C.byref(cfg), C.byref(inp), C.byref(out))
    assert rc == 0, rc
    return o


def forward_backward(args, grad_loss=1.0):
    a, cfg, inp = _prep(args)
    o = _alloc_out(cfg)
    gd, gT = _alloc_grads(a)
    out = cabi.make_outputs(o["loss"], o["per_pixel"], o["argmin"], o["depth"])
    g = cabi.make_grads(gd, gT)
    rc = lib().md2_emu_forward_backward(C.byref(cfg), C.byref(inp), C.byref(out), C.byref(g),
                                        C.c_float(grad_loss))
    assert rc == 0, rc
    o["grad_disp"], o["grad_T"] = gd, gT
    return o


def backward(args, argmin, grad_loss=1.0):
    a, cfg, inp = _prep(args)
    gd, gT = _alloc_grads(a)
    g = cabi.make_grads(gd, gT)
    am = argmin.contiguous()
    rc = lib().md2_emu_backward(C.byref(cfg), C.byref(inp), C.c_void_p(am.data_ptr()), C.c_float(grad_loss),
                                C.byref(g))
    assert rc == 0, rc
    return dict(grad_disp=gd, grad_T=gT)


def debug_warp(args, scale, source):
    a, cfg, inp = _prep(args)
    coords = torch.zeros(cfg.B, 2, cfg.H, cfg.W)
    warped = torch.zeros(cfg.B, 3, cfg.H, cfg.W)
    rc = lib().md2_emu_debug_warp(C.byref(cfg), C.byref(inp), scale, source, C.c_void_p(coords.data_ptr()),
                                  C.c_void_p(warped.data_ptr()))
    assert rc == 0, rc
    return coords, warped


def pose_forward(aa, tr, invert):
    aa, tr = aa.detach().contiguous().view(-1, 3), tr.detach().contiguous().view(-1, 3)
    M = torch.zeros(aa.shape[0], 4, 4)
    lib().md2_emu_pose_forward(aa.shape[0], C.c_void_p(aa.data_ptr()), C.c_void_p(tr.data_ptr()), int(invert),
                               C.c_void_p(M.data_ptr()))
    return M


def pose_backward(aa, tr, invert, gM):
    aa, tr = aa.detach().contiguous().view(-1, 3), tr.detach().contiguous().view(-1, 3)
    gM = gM.contiguous()
    ga, gt = torch.zeros_like(aa), torch.zeros_like(tr)
    lib().md2_emu_pose_backward(aa.shape[0], C.c_void_p(aa.data_ptr()), C.c_void_p(tr.data_ptr()), int(invert),
                                C.c_void_p(gM.data_ptr()), C.c_void_p(ga.data_ptr()), C.c_void_p(gt.data_ptr()))
    return ga, gt
