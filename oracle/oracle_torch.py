"""ORACLE (test infrastructure, not product code).

A restatement of the reference's view-synthesis training loss in plain PyTorch
ops, written from the maths of SURVEY.md Appendix A.  It exists to check the
CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
never does (it raises when its CUDA extension is missing).

Pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4, 8c), so the oracle is pinned against outputs of the
reference itself, run in the build container by ``tests/golden/make_golden.py``
and committed under ``tests/golden/*.npz`` (``tests/test_oracle_golden.py``).

Every function cites the reference lines it follows (paths relative to
/root/reference).  The same ATen ops are used as the reference uses, in the same
order, so that on one device and dtype the oracle is bit-identical to it; it
runs on CPU (fp32 / fp64) and, inside the GPU parity tests, on the B200.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- warp
def upsample_disp(disp, H, W):
    """model_layer/warp.py:18-20 as called from model_tool/processor.py:142
    (bilinear, align_corners=False)."""
    return F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)


def disp_to_depth(disp, min_depth, max_depth):
    """model_layer/warp.py:29-39: sigmoid disparity -> (scaled disparity, depth)."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def pixel_rays_grid(B, H, W, dtype, device):
    """model_layer/warp.py:207-234: homogeneous pixel coordinates [B,3,H*W],
    x fastest (meshgrid indexing 'xy'), rows (x, y, 1)."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype, device=device),
                            torch.arange(W, dtype=dtype, device=device), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1),
                       torch.ones(H * W, dtype=dtype, device=device)], 0)
    return pix[None].repeat(B, 1, 1)


def backproject(depth, inv_K, pix):
    """model_layer/warp.py:237-246: cam = depth * (inv_K[:3,:3] @ pix), then a row of ones."""
    B = depth.shape[0]
    rays = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.view(B, 1, -1) * rays
    ones = torch.ones(B, 1, cam.shape[-1], dtype=cam.dtype, device=cam.device)
    return torch.cat([cam, ones], 1)


def project(cam, K, T, H, W, eps=1e-7):
    """model_layer/warp.py:259-269: P=(K@T)[:3]; perspective divide; normalise to [-1,1]."""
    B = cam.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    xyz = torch.matmul(P, cam)
    uv = xyz[:, :2, :] / (xyz[:, 2, :].unsqueeze(1) + eps)
    uv = uv.view(B, 2, H, W).permute(0, 2, 3, 1)
    uv[..., 0] /= W - 1
    uv[..., 1] /= H - 1
    return (uv - 0.5) * 2


def sample_border(img, grid):
    """model_layer/warp.py:12-14 as called from processor.py:161."""
    return F.grid_sample(img, grid, padding_mode="border", align_corners=True)


# --------------------------------------------------------------------------- pose
def pose_matrix(axisangle, translation, invert=False):
    """model_layer/warp.py:43-153 (vector2translation, angle2rotation, param2matrix):
    Rodrigues rotation with axis = aa/(|aa|+1e-5); M = T@R, or R^T @ T(-t) when inverted."""
    n = axisangle.shape[0]
    dt, dev = axisangle.dtype, axisangle.device
    angle = torch.linalg.norm(axisangle, ord=2, dim=2, keepdim=True)
    axis = axisangle / (angle + 1e-5)
    c, s = torch.cos(angle), torch.sin(angle)
    C = 1 - c
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * s, y * s, z * s
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    R = torch.zeros(n, 4, 4, dtype=dt, device=dev)
    R[:, 0, 0] = torch.squeeze(x * xC + c)
    R[:, 0, 1] = torch.squeeze(xyC - zs)
    R[:, 0, 2] = torch.squeeze(zxC + ys)
    R[:, 1, 0] = torch.squeeze(xyC + zs)
    R[:, 1, 1] = torch.squeeze(y * yC + c)
    R[:, 1, 2] = torch.squeeze(yzC - xs)
    R[:, 2, 0] = torch.squeeze(zxC - ys)
    R[:, 2, 1] = torch.squeeze(yzC + xs)
    R[:, 2, 2] = torch.squeeze(z * zC + c)
    R[:, 3, 3] = 1
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    Tm = torch.zeros(n, 4, 4, dtype=dt, device=dev)
    Tm[:, 0, 0] = 1
    Tm[:, 1, 1] = 1
    Tm[:, 2, 2] = 1
    Tm[:, 3, 3] = 1
    Tm[:, :3, 3, None] = t.contiguous().view(-1, 3, 1)
    return torch.matmul(R, Tm) if invert else torch.matmul(Tm, R)


# --------------------------------------------------------------------------- loss
def ssim_dissimilarity(x, y):
    """model_loss/model_loss.py:28-41: clamp((1-SSIM)/2, 0, 1) with 3x3 box statistics
    over a reflection-padded border, C1=0.01^2, C2=0.03^2."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sig_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sig_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sig_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sig_x + sig_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def photometric_error(pred, target):
    """model_loss/model_loss.py:97-103: 0.85*mean_c(SSIM) + 0.15*mean_c|target-pred|."""
    l1 = torch.abs(target - pred).mean(1, True)
    ss = ssim_dissimilarity(pred, target).mean(1, True)
    return 0.85 * ss + 0.15 * l1


def smoothness(disp, color):
    """model_loss/model_loss.py:77-88,112-116: mean-normalised edge-aware smoothness."""
    mean_disp = disp.mean(2, True).mean(3, True)
    n = disp / (mean_disp + 1e-7)
    gdx = torch.abs(n[:, :, :, :-1] - n[:, :, :, 1:])
    gdy = torch.abs(n[:, :, :-1, :] - n[:, :, 1:, :])
    gix = torch.mean(torch.abs(color[:, :, :, :-1] - color[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(color[:, :, :-1, :] - color[:, :, 1:, :]), 1, keepdim=True)
    gdx = gdx * torch.exp(-gix)
    gdy = gdy * torch.exp(-giy)
    return gdx.mean() + gdy.mean()


def view_synthesis_loss(target, sources, disps, color_pyr, K, inv_K, Ts, *,
                        min_depth=0.1, max_depth=100.0, disp_smoothness=1e-3,
                        automask=True, noise=None, taps=False, posecnn=None):
    """model_tool/processor.py:139-218 (image2warping followed by compute_loss) for
    one batch, as a function of explicit tensors.

    target [B,3,H,W]; sources: S x [B,3,H,W]; disps / color_pyr: per scale
    [B,1,h,w] / [B,3,h,w]; K, inv_K [B,4,4]; Ts: S x [B,4,4].
    noise: per-scale [B,S,H,W] standard-normal draws, or None to draw them the way
    the reference does (host torch.randn per scale, processor.py:195).
    Returns a dict: loss (0-dim), depth[s], per_pixel[s] [B,H,W], argmin[s] [B,H,W]
    (+ warped / grid / rep / ident taps when ``taps``).
    posecnn: None, or per source (axisangle [B,1,3], translation [B,1,3], invert): the pose_type
    "posecnn" branch (processor.py:153-157), where the pose of each scale is built from the
    translation scaled by that scale's mean inverse depth (``Ts`` is then ignored)."""
    B, _, H, W = target.shape
    S = len(sources)
    pix = pixel_rays_grid(B, H, W, target.dtype, target.device)
    out = {"depth": [], "per_pixel": [], "argmin": [], "warped": [], "grid": [],
           "rep": [], "ident": [], "smooth": []}
    total = 0
    for s, disp in enumerate(disps):
        up = upsample_disp(disp, H, W)
        _, depth = disp_to_depth(up, min_depth, max_depth)
        out["depth"].append(depth)
        reps, warps, grids = [], [], []
        for f in range(S):
            cam = backproject(depth, inv_K, pix)
            if posecnn is not None:
                aa, tr, inv = posecnn[f]
                mean_inv_depth = (1 / depth).mean(3, True).mean(2, True)
                T = pose_matrix(aa, tr * mean_inv_depth[:, 0], invert=inv)
            else:
                T = Ts[f]
            grid = project(cam, K, T, H, W)
            w = sample_border(sources[f], grid)
            reps.append(photometric_error(w, target))
            if taps:
                warps.append(w)
                grids.append(grid)
        rep = torch.cat(reps, 1)
        if automask:
            ident = torch.cat([photometric_error(src, target) for src in sources], 1)
            if noise is None:
                nz = torch.randn(ident.shape).to(ident.device)
            else:
                nz = noise[s]
            ident = ident + 0.00001 * nz.to(ident.dtype)
            comb = torch.cat((ident, rep), dim=1)
        else:
            ident = None
            comb = rep
        if comb.shape[1] == 1:
            per_px, idx = comb[:, 0], torch.zeros_like(comb[:, 0], dtype=torch.long)
        else:
            per_px, idx = torch.min(comb, dim=1)
        sm = smoothness(disp, color_pyr[s])
        total = total + per_px.mean() + disp_smoothness * sm / (2 ** s)
        out["per_pixel"].append(per_px)
        out["argmin"].append(idx)
        out["smooth"].append(sm)
        if taps:
            out["warped"].append(warps)
            out["grid"].append(grids)
            out["rep"].append(rep)
            out["ident"].append(ident)
    out["loss"] = total / len(disps)
    return out


def loss_and_grads(target, sources, disps, color_pyr, K, inv_K, Ts, grad_T_mask=None, **kw):
    """Forward + autograd backward (model_train.py:67-68).  ``disps`` and the entries
    of ``Ts`` that need a gradient must be leaf tensors with requires_grad=True.
    Returns the forward dict plus grad_disp[s] and grad_T[f] (None where T is data)."""
    out = view_synthesis_loss(target, sources, disps, color_pyr, K, inv_K, Ts, **kw)
    wrt = list(disps) + [T for T in Ts if T.requires_grad]
    grads = torch.autograd.grad(out["loss"], wrt, allow_unused=True)
    out["grad_disp"] = list(grads[:len(disps)])
    it = iter(grads[len(disps):])
    out["grad_T"] = [next(it) if T.requires_grad else None for T in Ts]
    return out


def depth_metrics(depth, gt, crop=(153, 371, 44, 1197)):
    """model_loss/model_metric.py:70-106 (compute_depth_metric) with :43-63 (compute_depth_error,
    lib="torch"): up-sample the prediction to the ground-truth size, clamp, mask gt > 0 inside the crop,
    median scaling, clamp, then (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3)."""
    pred = torch.clamp(F.interpolate(depth, list(gt.shape[2:]), mode="bilinear", align_corners=False), 1e-3, 80)
    pred = pred.detach()
    mask = gt > 0
    crop_mask = torch.zeros_like(mask)
    crop_mask[:, :, crop[0]:crop[1], crop[2]:crop[3]] = 1
    mask = mask * crop_mask
    g = gt[mask]
    p = pred[mask]
    p = p * (torch.median(g) / torch.median(p))
    p = torch.clamp(p, min=1e-3, max=80)
    thr = torch.maximum(g / p, p / g)
    a1 = (thr < 1.25).float().mean()
    a2 = (thr < 1.25 ** 2).float().mean()
    a3 = (thr < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((g - p) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(g) - torch.log(p)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(g - p) / g)
    sq_rel = torch.mean((g - p) ** 2 / g)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
