"""First-call self-check of the cuBLAS rounding table.

The reference builds its rays and projections with torch.matmul (model_layer/warp.py:238,260-261), and which
rounding sequence cuBLAS applies to those 3- and 4-term dot products depends on the batch size and on kernel-
selection thresholds of the installed torch / cuBLAS (csrc/md2_host.h matmul_mode(): measured on torch 2.11 /
cuBLAS 12.8).  A different build could silently move a threshold: the loss would still be within tolerance, but no
longer bit-identical.  ``matmul_rounding`` runs the symbol-level kernels of this package - which use the same
table as the fused kernel - against torch.matmul on the device for one (batch, height, width) and reports whether
they agree bit for bit; ``compute`` calls it once per shape and warns when they do not.
"""
from __future__ import annotations

import warnings

import torch

_seen = {}


def matmul_rounding(B, H, W, device="cuda"):
    """-> dict(rays_bit_exact, projection_bit_exact, shape).  Cached per (B, H, W, device index)."""
    dev = torch.device(device)
    key = (int(B), int(H), int(W), dev.index if dev.index is not None else torch.cuda.current_device())
    if key in _seen:
        return _seen[key]
    from . import modules as M
    from .synthetic import make_intrinsics
    with torch.no_grad():
        g = torch.Generator().manual_seed(1234)
        K, inv_K = (t.to(dev) for t in make_intrinsics(B, H, W))
        depth = (0.1 + 99.9 * torch.rand(B, 1, H, W, generator=g)).to(dev)
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, :] += 0.05 * torch.randn(B, 3, 4, generator=g)
        T = T.to(dev)
        # the reference's ops (warp.py:237-246, 259-263)
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                                torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=dev)], 0)[None].repeat(B, 1, 1)
        cam_ref = depth.view(B, 1, -1) * torch.matmul(inv_K[:, :3, :3], pix)
        cam_ref = torch.cat([cam_ref, torch.ones(B, 1, H * W, device=dev)], 1)
        xyz_ref = torch.matmul(torch.matmul(K, T)[:, :3, :], cam_ref)
        uv_ref = xyz_ref[:, :2, :] / (xyz_ref[:, 2, :].unsqueeze(1) + 1e-7)
        uv_ref = uv_ref.view(B, 2, H, W).permute(0, 2, 3, 1).clone()
        uv_ref[..., 0] /= W - 1
        uv_ref[..., 1] /= H - 1
        grid_ref = (uv_ref - 0.5) * 2
        cam = M.Depth2PointCloud(B, H, W)(depth, inv_K)
        grid = M.PointCloud2Pixel(B, H, W)(cam_ref, K, T)
        res = dict(rays_bit_exact=bool(torch.equal(cam, cam_ref)), projection_bit_exact=bool(torch.equal(grid, grid_ref)),
                   shape=(B, H, W), torch=torch.__version__)
    _seen[key] = res
    return res


def warn_if_not_replicated(B, H, W, device):
    r = matmul_rounding(B, H, W, device)
    if not (r["rays_bit_exact"] and r["projection_bit_exact"]):
        warnings.warn(
            f"md2_b200: torch.matmul on this build ({r['torch']}) rounds the ray / projection products of shape "
            f"B={B}, H={H}, W={W} differently from the table in csrc/md2_host.h (rays bit-exact: {r['rays_bit_exact']}, "
            f"projection bit-exact: {r['projection_bit_exact']}); results stay within the 1e-5 / 1e-4 tolerances of "
            f"DESIGN.md 3 but are not bit-identical to the reference's PyTorch path.", RuntimeWarning, stacklevel=3)
    return r
