"""Exploratory parity / timing probe on the B200 (development tool, not a test).

Prints per-stage agreement of the CUDA path with the oracle run on the same GPU in fp32
and fp64 (bit-exact rates, max errors, argmin flips), then a few timings.
Usage: python tools/gpu_probe.py [--quick]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import md2_b200.cabi as cabi  # noqa: E402
import md2_b200.synthetic as syn  # noqa: E402
from oracle import oracle_torch as O  # noqa: E402

dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def synth_args(B, H, W, frame_ids, automask, kind, seed, num_scales=4):
    inputs, outputs = syn.make_batch(B, H, W, frame_ids, num_scales, seed, kind, requires_grad=False)
    srcs = frame_ids[1:]
    Ts = []
    for f in srcs:
        if f == "s":
            Ts.append(inputs["stereo"].to(dev))
        else:
            Ts.append(O.pose_matrix(outputs[("axisangle", f)].to(dev), outputs[("translation", f)].to(dev),
                                    invert=(f < 0)).detach())
    g = lambda t: t.to(dev)
    return dict(target=g(inputs[("color", 0, 0)]), sources=[g(inputs[("color", f, 0)]) for f in srcs],
                disps=[g(outputs[("disp", s)]) for s in range(num_scales)],
                color_pyr=[g(inputs[("color", 0, s)]) for s in range(num_scales)],
                K=g(inputs[("K", 0)]), inv_K=g(inputs[("inv_K", 0)]), Ts=Ts, automask=automask,
                noise=[g(n) for n in syn.make_noise(B, len(srcs), H, W, num_scales, seed)] if automask else None)


def to64(args):
    cv = lambda v: [t.double() for t in v] if isinstance(v, list) else (v.double() if torch.is_tensor(v) else v)
    return {k: cv(v) for k, v in args.items()}


def with_grad(args):
    a = dict(args)
    a["disps"] = [d.clone().requires_grad_(True) for d in args["disps"]]
    a["Ts"] = [T.clone().requires_grad_(True) for T in args["Ts"]]
    return a


def biteq(a, b):
    return float((a == b).float().mean())


def nrel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def mrel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def parity(cl, name, args):
    B, _, H, W = args["target"].shape
    S = len(args["sources"])
    r32 = O.loss_and_grads(**with_grad(args), taps=True)
    r64 = O.loss_and_grads(**with_grad(to64(args)), taps=True)
    out = cl.forward_backward(args)
    rep = {"case": name, "loss": [float(out["loss"]), float(r32["loss"]), float(r64["loss"])]}
    for s in range(len(args["disps"])):
        d = {}
        d["depth_biteq"] = biteq(out["depth"][s], r32["depth"][s])
        for f in range(S):
            coords, warped = cl.debug_warp(args, s, f)
            g = r32["grid"][s][f]
            ix = ((g[..., 0] + 1) / 2) * (W - 1)
            iy = ((g[..., 1] + 1) / 2) * (H - 1)
            d[f"ix_biteq_f{f}"] = biteq(coords[:, 0], ix)
            d[f"iy_biteq_f{f}"] = biteq(coords[:, 1], iy)
            d[f"ix_maxabs_f{f}"] = float((coords[:, 0] - ix).abs().max())
            d[f"warp_biteq_f{f}"] = biteq(warped, r32["warped"][s][f])
            d[f"warp_maxabs_f{f}"] = float((warped - r32["warped"][s][f]).abs().max())
        pp, p32, p64 = out["per_pixel"][s], r32["per_pixel"][s], r64["per_pixel"][s]
        d["px_biteq"] = biteq(pp, p32)
        d["px_maxrel_vs32"] = mrel(pp, p32)
        d["px_maxabs_vs32"] = float((pp - p32).abs().max())
        d["px_maxabs_ours_vs64"] = float((pp.double() - p64).abs().max())
        d["px_maxabs_ref32_vs64"] = float((p32.double() - p64).abs().max())
        mism = out["argmin"][s].long() != r32["argmin"][s]
        d["argmin_flips_vs32"] = int(mism.sum())
        d["argmin_flips_ref32_vs64"] = int((r32["argmin"][s] != r64["argmin"][s]).sum())
        d["gdisp_nrel_vs32"] = nrel(out["grad_disp"][s], r32["grad_disp"][s])
        d["gdisp_nrel_vs64"] = nrel(out["grad_disp"][s], r64["grad_disp"][s])
        d["gdisp_nrel_ref32_vs64"] = nrel(r32["grad_disp"][s], r64["grad_disp"][s])
        rep[f"scale{s}"] = d
    for f in range(S):
        rep[f"gT{f}"] = [nrel(out["grad_T"][f], r32["grad_T"][f]), nrel(out["grad_T"][f], r64["grad_T"][f]),
                         nrel(r32["grad_T"][f], r64["grad_T"][f])]
    return rep


def time_fn(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    quick = "--quick" in sys.argv
    cl = cabi.CLoss()
    print(torch.cuda.get_device_name(0), cl.lib.md2_version())
    reps = []
    reps.append(parity(cl, "smooth_B2_192x640", synth_args(2, 192, 640, [0, -1, 1], True, "smooth", 0)))
    reps.append(parity(cl, "iid_B1_192x640", synth_args(1, 192, 640, [0, -1, 1], True, "iid", 1)))
    if not quick:
        reps.append(parity(cl, "stereo_B1_96x320_S3", synth_args(1, 96, 320, [0, -1, 1, "s"], True, "smooth", 2)))
        reps.append(parity(cl, "nomask_B1_64x96", synth_args(1, 64, 96, [0, -1, 1], False, "smooth", 3)))
    for r in reps:
        print(json.dumps(r, indent=1))
    # timings at the benchmark configuration
    args = synth_args(12, 192, 640, [0, -1, 1], True, "iid", 0)
    t_fused = time_fn(lambda: cl.forward_backward(args))
    t_fwd = time_fn(lambda: cl.forward(args))
    am = cl.forward(args)["argmin"]
    t_bwd = time_fn(lambda: cl.backward(args, am, 1.0))

    def ref_step():
        a = with_grad(args)
        out = O.view_synthesis_loss(**a)
        out["loss"].backward()
    t_ref = time_fn(ref_step, iters=5, warm=2)
    px = 12 * 2 * 4 * 192 * 640
    algo = 12 * 192 * 640 * (183.8125 + 112 * 2)
    print(json.dumps({"timing_ms": {"fused_fwd_bwd": t_fused, "fwd_only": t_fwd, "bwd_standalone": t_bwd,
                                    "oracle_eager_gpu_fwd_bwd": t_ref},
                      "warped_px_per_s_fused": px / (t_fused * 1e-3),
                      "roofline_frac_fused": algo / (t_fused * 1e-3) / 6539.2e9}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(reps, f, indent=1)


if __name__ == "__main__":
    main()
