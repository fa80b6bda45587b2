"""Achieved HBM bandwidth of the channels-last network operators: md2_b200.modules.ReflectionPad2d (csrc/md2_pad.cu)
on the fourteen Conv3x3 inputs of the ResNet-18 depth decoder, and md2_b200.modules.MaxPool2d (csrc/md2_pool.cu) on the
stems of the depth and pose encoders, at batch 12, 192x640 - forward and backward, each list replayed from one CUDA
graph (the launches are shorter than their Python calls) - next to ATen's operators on the same tensors.
Two JSON lines.  python tools/bench_nhwc_ops.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import md2_b200.modules as M  # noqa: E402

DEV = "cuda:0"
B = 12
# (channels, height, width) of every Conv3x3 input: upconv(i,0), upconv(i,1) for i = 4..0, then the four dispconvs
SHAPES = [(512, 6, 20), (512, 12, 40), (256, 12, 40), (256, 24, 80), (128, 24, 80), (128, 48, 160), (64, 48, 160),
          (96, 96, 320), (32, 96, 320), (16, 192, 640), (16, 192, 640), (32, 96, 320), (64, 48, 160), (128, 24, 80)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run(name_md2, make_md2, make_aten, xs, gos, bytes_fwd, bytes_bwd, workload, note):
    res = {}
    for name, op in (("md2", make_md2()), ("aten", make_aten())):
        with torch.no_grad():
            fwd_ms = timed(lambda: [op(x) for x in xs])

        def both():   # backward runs on the stream of its forward: capture the pair, subtract the forward time
            torch.autograd.grad([op(x) for x in xs], xs, gos)
        bwd_ms = timed(both) - fwd_ms
        res[name] = {"forward_ms": round(fwd_ms, 4), "backward_ms": round(bwd_ms, 4)}
    peak = 6539.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for k, b in (("forward", bytes_fwd), ("backward", bytes_bwd)):
        gbs = b / (res["md2"][k + "_ms"] * 1e-3) / 1e9
        res["md2"][k + "_GBps"] = round(gbs, 1)
        res["md2"][k + "_roofline_frac"] = round(gbs / peak, 4)
    print(json.dumps({"op": name_md2, "workload": workload, "algorithmic_MB_forward": round(bytes_fwd / 1e6, 1),
                      "algorithmic_MB_backward": round(bytes_bwd / 1e6, 1), "peak_GBps": peak, **res, "note": note}), flush=True)


def main():
    # ---- max-pool of the encoder stems: depth encoder on 12 images, pose encoder on 2 x 12 image pairs
    pxs = [torch.randn(n, 64, 96, 320, device=DEV).clamp_min(0).contiguous(memory_format=torch.channels_last).requires_grad_(True)
           for n in (B, 2 * B)]
    pgos = [torch.randn(x.shape[0], 64, 48, 160, device=DEV).contiguous(memory_format=torch.channels_last) for x in pxs]
    pin, pout = sum(x.numel() for x in pxs), sum(g.numel() for g in pgos)
    run("md2_b200.modules.MaxPool2d", lambda: M.MaxPool2d(3, 2, 1), lambda: torch.nn.MaxPool2d(3, 2, 1), pxs, pgos,
        4 * pin + 5 * pout, 5 * pout + 4 * pin,
        f"MaxPool2d(3, 2, 1) on the encoder stems [{B}|{2 * B}, 64, 96, 320], channels-last, fp32",
        "aten = nn.MaxPool2d on the same channels-last tensors (int64 indices, atomics + zero fill in backward)")
    del pxs, pgos
    torch.cuda.empty_cache()
    pad_section()


def pad_section():
    xs = [torch.randn(B, c, h, w, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for c, h, w in SHAPES]
    gos = [torch.randn(B, c, h + 2, w + 2, device=DEV).contiguous(memory_format=torch.channels_last) for c, h, w in SHAPES]
    n_in = sum(x.numel() for x in xs)
    n_out = sum(g.numel() for g in gos)
    run("md2_b200.modules.ReflectionPad2d", lambda: M.ReflectionPad2d(1), lambda: torch.nn.ReflectionPad2d(1), xs, gos,
        4 * (n_in + n_out), 4 * (n_in + n_out),
        f"ReflectionPad2d(1) on the 14 Conv3x3 inputs of the depth decoder, batch {B}, 192x640, channels-last, fp32",
        "aten = nn.ReflectionPad2d on the same channels-last tensors (converts to NCHW and back inside the operator)")


if __name__ == "__main__":
    main()
