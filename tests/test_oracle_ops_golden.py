"""Pins the oracle's per-operator functions against the reference's own symbols (tests/golden/ops.npz,
written by tests/golden/make_golden_ops.py from model_layer/warp.py and model_loss/model_loss.py)."""
import os

import numpy as np
import torch

from oracle import oracle_torch as O
from helpers import GOLDEN_DIR


def _load(name="ops.npz"):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _eq(a, b, tol=0.0):
    a, b = a.detach(), b.detach()
    if tol == 0.0:
        assert torch.equal(a, b), float((a - b).abs().max())
    else:
        assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_upsample_and_disp2depth():
    z = _load()
    x = z["up_in"].clone().requires_grad_(True)
    up = O.upsample_disp(x, *z["up_out"].shape[2:])
    _eq(up, z["up_out"])
    _eq(torch.autograd.grad((up * z["up_cot"]).sum(), x)[0], z["up_grad"], 1e-6)
    d = z["d2d_in"].clone().requires_grad_(True)
    scaled, depth = O.disp_to_depth(d, 0.1, 100.0)
    _eq(scaled, z["d2d_scaled"])
    _eq(depth, z["d2d_depth"])
    _eq(torch.autograd.grad((scaled * z["d2d_cot0"]).sum() + (depth * z["d2d_cot1"]).sum(), d)[0], z["d2d_grad"], 1e-6)


def test_backproject_project_sample():
    z = _load()
    B, _, H, W = z["bp_depth"].shape
    pix = O.pixel_rays_grid(B, H, W, torch.float32, "cpu")
    dep = z["bp_depth"].clone().requires_grad_(True)
    cam = O.backproject(dep, z["bp_invK"], pix)
    _eq(cam, z["bp_cam"])
    _eq(torch.autograd.grad((cam * z["bp_cot"]).sum(), dep)[0], z["bp_grad"], 1e-6)
    c = z["pj_cam"].clone().requires_grad_(True)
    T = z["pj_T"].clone().requires_grad_(True)
    grid = O.project(c, z["pj_K"], T, H, W)
    _eq(grid, z["pj_grid"])
    gc, gT = torch.autograd.grad((grid * z["pj_cot"]).sum(), [c, T])
    _eq(gc, z["pj_grad_cam"], 1e-6)
    _eq(gT, z["pj_grad_T"], 1e-5)
    g = z["gs_grid"].clone().requires_grad_(True)
    out = O.sample_border(z["gs_img"], g)
    _eq(out, z["gs_out"])
    _eq(torch.autograd.grad((out * z["gs_cot"]).sum(), g)[0], z["gs_grad"], 1e-6)


def test_reprojection_and_smoothness():
    z = _load()
    p = z["rp_pred"].clone().requires_grad_(True)
    rep = O.photometric_error(p, z["rp_target"])
    _eq(rep, z["rp_out"])
    _eq(torch.autograd.grad((rep * z["rp_cot"]).sum(), p)[0], z["rp_grad"], 1e-6)
    d = z["sm_disp"].clone().requires_grad_(True)
    sm = O.smoothness(d, z["sm_color"])
    _eq(sm, z["sm_out"])
    _eq(torch.autograd.grad(sm, d)[0], z["sm_grad"], 1e-6)


def posecnn_args(z, device="cpu"):
    t = lambda k: z[k].to(device)
    leaf = lambda k: z[k].to(device).clone().requires_grad_(True)
    disps = [leaf(f"disp{s}") for s in range(4)]
    R = {f: leaf(f"R{f}") for f in (-1, 1)}
    T = {f: leaf(f"T{f}") for f in (-1, 1)}
    return dict(target=t("color0"), sources=[t("color-1"), t("color1")], disps=disps,
                color_pyr=[t(f"color_pyr{s}") for s in range(4)], K=t("K"), inv_K=t("inv_K"),
                noise=[t(f"noise{s}") for s in range(4)], R=R, T=T)


def test_posecnn_branch():
    """processor.py:153-157 (pose rebuilt per scale from the mean inverse depth) against the reference's run."""
    z = _load("posecnn.npz")
    a = posecnn_args(z)
    out = O.view_synthesis_loss(a["target"], a["sources"], a["disps"], a["color_pyr"], a["K"], a["inv_K"], None,
                                noise=a["noise"], taps=True,
                                posecnn=[(a["R"][f][:, 0], a["T"][f][:, 0], f < 0) for f in (-1, 1)])
    _eq(out["loss"], z["loss"])
    for s in range(4):
        _eq(out["warped"][s][1], z[f"warp{s}"])
    wrt = a["disps"] + [a["R"][-1], a["R"][1], a["T"][-1], a["T"][1]]
    grads = torch.autograd.grad(out["loss"], wrt)
    for s in range(4):
        _eq(grads[s], z[f"grad_disp{s}"], 1e-6)
    for g, k in zip(grads[4:], ["grad_R-1", "grad_R1", "grad_T-1", "grad_T1"]):
        _eq(g, z[k], 1e-5)


def metrics_golden(device="cpu"):
    z = np.load(os.path.join(GOLDEN_DIR, "metrics.npz"))
    depth = torch.from_numpy(z["depth"])
    gt = torch.zeros(depth.shape[0] * 375 * 1242)
    gt[torch.from_numpy(z["gt_idx"]).long()] = torch.from_numpy(z["gt_val"])
    return depth.to(device), gt.view(depth.shape[0], 1, 375, 1242).to(device), z["metrics"], int(z["n"])


def test_depth_metrics():
    depth, gt, ref, _ = metrics_golden()
    got = O.depth_metrics(depth, gt)
    for a, b in zip(got, ref):
        assert float(a) == float(np.float32(b)) or abs(float(a) - b) <= 1e-7 * abs(b), (float(a), b)
