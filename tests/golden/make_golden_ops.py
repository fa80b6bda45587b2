"""Golden vectors for the symbol-level operators (SURVEY.md 8b "L1"), produced by the UNMODIFIED reference
symbols of model_layer/warp.py and model_loss/model_loss.py on CPU fp32.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_ops.py
Writes tests/golden/ops.npz: for every operator its inputs, its output, a fixed cotangent and the gradients
autograd returns for that cotangent.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MD2_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402

for _m in ["matplotlib", "matplotlib.pyplot", "albumentations", "albumentations.pytorch",
           "albumentations.pytorch.transforms", "albumentations.augmentations",
           "albumentations.augmentations.transforms", "skimage", "skimage.transform", "cv2"]:
    sys.modules[_m] = MagicMock()
sys.modules["albumentations"].__version__ = "0.5.2"

from model_layer.warp import (Depth2PointCloud, PointCloud2Pixel, disparity2depth, grid_sample,  # noqa: E402
                              interpolate)
from model_loss.model_loss import ReprojectionLoss, SmoothLoss  # noqa: E402
from model_tool.processor import compute  # noqa: E402

import md2_b200.synthetic as syn  # noqa: E402

B, H, W = 2, 24, 32


def main():
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: torch.rand(*s, generator=g)
    d = {}

    disp_lo = rnd(B, 1, H // 2, W // 2).requires_grad_(True)
    up = interpolate(disp_lo, H, W, "bilinear", False)
    cot = torch.randn(up.shape, generator=g)
    d.update(up_in=disp_lo, up_out=up, up_cot=cot, up_grad=torch.autograd.grad((up * cot).sum(), disp_lo)[0])

    disp = rnd(B, 1, H, W).requires_grad_(True)
    scaled, depth = disparity2depth(disp, 0.1, 100.0)
    c0, c1 = torch.randn(scaled.shape, generator=g), torch.randn(depth.shape, generator=g) * 1e-2
    d.update(d2d_in=disp, d2d_scaled=scaled, d2d_depth=depth, d2d_cot0=c0, d2d_cot1=c1,
             d2d_grad=torch.autograd.grad((scaled * c0).sum() + (depth * c1).sum(), disp)[0])

    K4, inv_K = syn.make_intrinsics(B, H, W, "monodepth2")
    depth_in = (0.5 + 5 * rnd(B, 1, H, W)).requires_grad_(True)
    cam = Depth2PointCloud(B, H, W)(depth_in, inv_K)
    cot = torch.randn(cam.shape, generator=g)
    d.update(bp_depth=depth_in, bp_invK=inv_K, bp_cam=cam, bp_cot=cot,
             bp_grad=torch.autograd.grad((cam * cot).sum(), depth_in)[0])

    T = torch.eye(4)[None].repeat(B, 1, 1)
    T[:, :3, :3] += 0.02 * torch.randn(B, 3, 3, generator=g)
    T[:, :3, 3] = 0.2 * torch.randn(B, 3, generator=g)
    T.requires_grad_(True)
    cam_in = cam.detach().clone().requires_grad_(True)
    grid = PointCloud2Pixel(B, H, W)(cam_in, K4, T)
    cot = torch.randn(grid.shape, generator=g)
    gc, gT = torch.autograd.grad((grid * cot).sum(), [cam_in, T])
    d.update(pj_cam=cam_in, pj_K=K4, pj_T=T, pj_grid=grid, pj_cot=cot, pj_grad_cam=gc, pj_grad_T=gT)

    img = rnd(B, 3, H, W)
    grid_in = (grid.detach() + 0.05 * torch.randn(grid.shape, generator=g)).requires_grad_(True)  # some out of range
    out = grid_sample(img, grid_in, "border", True)
    cot = torch.randn(out.shape, generator=g)
    d.update(gs_img=img, gs_grid=grid_in, gs_out=out, gs_cot=cot,
             gs_grad=torch.autograd.grad((out * cot).sum(), grid_in)[0])

    pred = rnd(B, 3, H, W).requires_grad_(True)
    tgt = (pred.detach() + 0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    rep = ReprojectionLoss()(pred, tgt)
    cot = torch.randn(rep.shape, generator=g)
    d.update(rp_pred=pred, rp_target=tgt, rp_out=rep, rp_cot=cot,
             rp_grad=torch.autograd.grad((rep * cot).sum(), pred)[0])

    sdisp = rnd(B, 1, H, W).requires_grad_(True)
    sm = SmoothLoss()(sdisp, img)
    d.update(sm_disp=sdisp, sm_color=img, sm_out=sm, sm_grad=torch.autograd.grad(sm, sdisp)[0])

    path = os.path.join(HERE, "ops.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in d.items()})
    print("ops.npz", os.path.getsize(path) // 1024, "KiB")


def posecnn_case(seed=9, Bp=2, Hp=32, Wp=64):
    """The reference's compute.image2warping + compute_loss with pose_type='posecnn' (processor.py:153-157)."""
    from types import SimpleNamespace
    torch.manual_seed(seed)
    frame_ids = [0, -1, 1]
    inputs, outputs = syn.make_batch(Bp, Hp, Wp, frame_ids, 4, seed, "smooth", "monodepth2")
    for f in frame_ids[1:]:
        outputs[("R", f, 0)] = outputs.pop(("axisangle", f)).detach()[:, None].clone().requires_grad_(True)
        outputs[("T", f, 0)] = outputs.pop(("translation", f)).detach()[:, None].clone().requires_grad_(True)
    opt = SimpleNamespace(frame_ids=frame_ids, scales=range(4), height=Hp, width=Wp, min_depth=0.1, max_depth=100.0,
                          pose_type="posecnn", pose_frames=2, use_automasking=True, disp_smoothness=1e-3, batch=Bp)
    setting = SimpleNamespace(inv_projection={0: Depth2PointCloud(Bp, Hp, Wp)},
                              for_projection={0: PointCloud2Pixel(Bp, Hp, Wp)},
                              loss={"reprojection": ReprojectionLoss(), "edge_aware": SmoothLoss()})
    c = compute(opt, "cpu")
    noise, orig = [], torch.randn

    def rec(*a, **k):
        r = orig(*a, **k)
        noise.append(r.clone())
        return r

    torch.randn = rec
    try:
        c.image2warping(inputs, outputs, setting)
        c.compute_loss(inputs, outputs, setting)
    finally:
        torch.randn = orig
    outputs["loss"].backward()
    d = {"loss": outputs["loss"].detach(), "K": inputs[("K", 0)], "inv_K": inputs[("inv_K", 0)]}
    for f in frame_ids:
        d[f"color{f}"] = inputs[("color", f, 0)]
    for f in frame_ids[1:]:
        d[f"R{f}"], d[f"T{f}"] = outputs[("R", f, 0)], outputs[("T", f, 0)]
        d[f"grad_R{f}"], d[f"grad_T{f}"] = outputs[("R", f, 0)].grad, outputs[("T", f, 0)].grad
    for s in range(4):
        d[f"disp{s}"], d[f"grad_disp{s}"] = outputs[("disp", s)], outputs[("disp", s)].grad
        d[f"color_pyr{s}"], d[f"noise{s}"] = inputs[("color", 0, s)], noise[s]
        d[f"warp{s}"] = outputs[("warp_color", 1, s)]
    path = os.path.join(HERE, "posecnn.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in d.items()})
    print("posecnn.npz", os.path.getsize(path) // 1024, "KiB, loss", float(d["loss"]))


def metrics_case(seed=13, Bm=2, Hm=48, Wm=160):
    """compute_depth_metric (model_loss/model_metric.py:70-106) on a LiDAR-like sparse ground truth."""
    from model_loss.model_metric import compute_depth_metric
    g = torch.Generator().manual_seed(seed)
    depth = 1.0 + 40 * torch.rand(Bm, 1, Hm, Wm, generator=g)
    depth[0, 0, :4, :4] = 200.0       # clamped to 80
    depth[1, 0, -4:, -4:] = 1e-5      # clamped to 1e-3
    gt = torch.zeros(Bm, 1, 375, 1242)
    hit = torch.rand(Bm, 1, 375, 1242, generator=g) < 0.02
    gt[hit] = (2.0 + 60 * torch.rand(int(hit.sum()), generator=g))
    m = compute_depth_metric({("depth", 0): gt}, {("depth", 0, 0): depth}, "torch")
    path = os.path.join(HERE, "metrics.npz")
    np.savez_compressed(path, depth=depth.numpy(), gt_idx=hit.flatten().nonzero()[:, 0].numpy().astype(np.int32),
                        gt_val=gt[hit].numpy(), metrics=np.array([float(v) for v in m], dtype=np.float64),
                        n=np.int64(int((hit[:, :, 153:371, 44:1197]).sum())))
    print("metrics.npz", os.path.getsize(path) // 1024, "KiB", [round(float(v), 5) for v in m])


if __name__ == "__main__":
    main()
    posecnn_case()
    metrics_case()
