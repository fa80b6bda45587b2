"""Synthetic KITTI-shaped batches for the view-synthesis loss path.

The reference has no dataset in this environment, so every
test and benchmark feeds the path with tensors of the layout its loaders emit
(/root/reference/model_loader/kitti_mono.py:195-215, kitti_stereo.py:239-256):

  inputs[("color", f, s)]  [B,3,H/2^s,W/2^s] fp32 in [0,1]     f in frame_ids, s in scales
  inputs[("K", 0)], inputs[("inv_K", 0)]  [B,4,4] fp32          inv_K = pinv(K) in fp32
  inputs["stereo"]         [B,4,4] identity with [0,3] = +-0.1  (only when "s" in frame_ids)
  outputs[("disp", s)]     [B,1,H/2^s,W/2^s] sigmoid range       (requires_grad)
  outputs[("c2c", f, 0)]   [B,4,4] pose matrices                 (requires_grad)

Everything is generated on the CPU with an explicit torch.Generator so that the
same seed gives the same batch on every machine, then moved to ``device``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def make_intrinsics(B, H, W, variant="monodepth2"):
    """K / inv_K as the reference loaders build them.

    variant "monodepth2": normalised KITTI intrinsics scaled by (W, H)
    (kitti_mono.py:62-65,203-206); "floor": the floor-divided stereo variant
    (kitti_stereo.py:239-241); "row1_width": KITTIMonoDataset_v2 multiplies row 1
    by the width (kitti_mono.py:326-327)."""
    K = np.array([[0.58, 0, 0.5, 0],
                  [0, 1.92, 0.5, 0],
                  [0, 0, 1, 0],
                  [0, 0, 0, 1]], dtype=np.float32)
    Kc = K.copy()
    if variant == "monodepth2":
        Kc[0, :] = Kc[0, :] * W
        Kc[1, :] = Kc[1, :] * H
    elif variant == "floor":
        Kc[0, :] = Kc[0, :] * W // 1
        Kc[1, :] = Kc[1, :] * H // 1
    elif variant == "row1_width":
        Kc[0, :] = Kc[0, :] * W // 1
        Kc[1, :] = Kc[1, :] * W // 1
    else:
        raise ValueError(variant)
    inv = np.linalg.pinv(Kc).astype(np.float32)
    Kt = torch.from_numpy(Kc)[None].repeat(B, 1, 1).contiguous()
    it = torch.from_numpy(inv)[None].repeat(B, 1, 1).contiguous()
    return Kt, it


def _smooth_field(gen, B, C, H, W, down):
    """U(0,1) at (H/down, W/down), bicubic-upsampled and clamped: photo-like,
    exercises the E[x^2]-mu^2 cancellation of SSIM."""
    h, w = max(H // down, 2), max(W // down, 2)
    lo = torch.rand(B, C, h, w, generator=gen)
    up = F.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    return up.clamp_(0.0, 1.0).contiguous()


def make_pose_params(B, n_src, seed):
    gen = torch.Generator().manual_seed(1000 + seed)
    aa = 0.01 * torch.randn(B, n_src, 1, 3, generator=gen)
    tr = 0.01 * torch.randn(B, n_src, 1, 3, generator=gen)
    return aa, tr


def make_batch(B, H, W, frame_ids=(0, -1, 1), num_scales=4, seed=0, kind="smooth",
               k_variant="monodepth2", device="cpu", pose_fn=None, requires_grad=True):
    """Returns (inputs, outputs) dicts shaped like the reference's
    (model_tool/processor.py:139-218 reads exactly these keys).

    kind "smooth": band-limited images / disparities (parity); "iid": i.i.d.
    U(0,1) (worst-case gather locality, throughput)."""
    gen = torch.Generator().manual_seed(seed)
    inputs, outputs = {}, {}
    for f in frame_ids:
        for s in range(num_scales):
            h, w = H >> s, W >> s
            if kind == "smooth":
                img = _smooth_field(gen, B, 3, h, w, 8 if s == 0 else max(8 >> s, 1))
            else:
                img = torch.rand(B, 3, h, w, generator=gen)
            inputs[("color", f, s)] = img.to(device)
    if kind == "smooth":
        # consecutive frames of one scene: sources are small perturbations of the target
        for f in frame_ids[1:]:
            t = inputs[("color", 0, 0)].cpu()
            shift = 2 if f == "s" else int(f) * 2
            moved = torch.roll(t, shifts=shift, dims=3)
            pert = 0.05 * (_smooth_field(gen, B, 3, H, W, 8) - 0.5)
            inputs[("color", f, 0)] = (moved + pert).clamp_(0, 1).contiguous().to(device)
    K, invK = make_intrinsics(B, H, W, k_variant)
    inputs[("K", 0)] = K.to(device)
    inputs[("inv_K", 0)] = invK.to(device)
    for s in range(num_scales):
        h, w = H >> s, W >> s
        if kind == "smooth":
            d = _smooth_field(gen, B, 1, h, w, max(16 >> s, 1))
        else:
            d = torch.rand(B, 1, h, w, generator=gen)
        d = d.to(device)
        d.requires_grad_(requires_grad)
        outputs[("disp", s)] = d
    srcs = [f for f in frame_ids[1:]]
    n_pose = len([f for f in srcs if f != "s"])
    aa, tr = make_pose_params(B, max(n_pose, 1), seed)
    i = 0
    for f in srcs:
        if f == "s":
            T = torch.eye(4)[None].repeat(B, 1, 1)
            T[:, 0, 3] = 0.1
            inputs["stereo"] = T.to(device)
        else:
            a = aa[:, i].clone().to(device).requires_grad_(requires_grad)
            t = tr[:, i].clone().to(device).requires_grad_(requires_grad)
            outputs[("axisangle", f)] = a
            outputs[("translation", f)] = t
            if pose_fn is not None:
                outputs[("c2c", f, 0)] = pose_fn(a, t, invert=(f < 0))
            i += 1
    return inputs, outputs


def make_noise(B, S, H, W, num_scales, seed, device="cpu"):
    """The N(0,1) draws of the auto-mask tie-breaker (processor.py:195), one
    [B,S,H,W] tensor per scale, so both sides of a parity test see the same noise."""
    gen = torch.Generator().manual_seed(7777 + seed)
    return [torch.randn(B, S, H, W, generator=gen).to(device) for _ in range(num_scales)]
