"""Shared helpers for the parity tests: golden-fixture loading and error metrics."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["mono_automask", "mono_iid", "stereo_automask", "mono_nomask", "single_nomask", "five_frames",
                "partial_tiles"]


def load_golden(name, device="cpu", dtype=torch.float32):
    """Returns (args, ref): args feed oracle.view_synthesis_loss / the fused op, ref holds
    the reference's recorded outputs (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = lambda k: torch.from_numpy(z[k]).to(device=device, dtype=dtype)
    S = len([k for k in z.files if k.startswith("source")])
    automask = bool(int(z["meta_automask"]))
    args = dict(
        target=t("target"),
        sources=[t(f"source{i}") for i in range(S)],
        disps=[t(f"disp{s}") for s in range(4)],
        color_pyr=[t(f"color_pyr{s}") for s in range(4)],
        K=t("K"), inv_K=t("inv_K"),
        Ts=[t(f"T{i}") for i in range(S)],
        automask=automask,
        noise=[t(f"noise{s}") for s in range(4)] if automask else None,
    )
    ref = {"loss": torch.from_numpy(z["loss"]),
           "depth": [torch.from_numpy(z[f"depth{s}"]) for s in range(4)],
           "grad_disp": [torch.from_numpy(z[f"grad_disp{s}"]) for s in range(4)],
           "grad_T": [torch.from_numpy(z[f"grad_T{i}"]) if f"grad_T{i}" in z.files else None
                      for i in range(S)]}
    if "per_pixel0" in z.files:
        ref["per_pixel"] = [torch.from_numpy(z[f"per_pixel{s}"]) for s in range(4)]
        ref["argmin"] = [torch.from_numpy(z[f"argmin{s}"]) for s in range(4)]
    return args, ref


def with_grad(args):
    a = dict(args)
    a["disps"] = [d.clone().requires_grad_(True) for d in args["disps"]]
    a["Ts"] = [T.clone().requires_grad_(True) for T in args["Ts"]]
    return a


def max_rel(a, b, floor=1e-12):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def norm_rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
