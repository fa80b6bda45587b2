"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

The reference is imported from /root/reference with sys.modules stubs for the
plotting / augmentation packages it imports at module top and never uses on
this path.  ``torch.randn`` and ``torch.min`` are wrapped only to
RECORD what the reference draws / computes (auto-mask noise, per-pixel minimum
and argmin are locals of compute.compute_loss, processor.py:195,204); the
wrapped calls return the original results unchanged.

Each case is written to tests/golden/<name>.npz: every input tensor, the
recorded noise, and the reference's loss, depth, per-pixel loss, argmin,
d loss / d disp_s and d loss / d T_f.
"""
import os
import sys
from types import SimpleNamespace
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
           "albumentations.augmentations.transforms", "skimage", "skimage.transform"]:
    sys.modules[_m] = MagicMock()
sys.modules["albumentations"].__version__ = "0.5.2"

from model_tool.processor import compute  # noqa: E402  (the reference)
from model_layer import Depth2PointCloud, PointCloud2Pixel, param2matrix  # noqa: E402
from model_loss import ReprojectionLoss, SmoothLoss  # noqa: E402

import md2_b200.synthetic as syn  # noqa: E402

CASES = {
    # name: (B, H, W, frame_ids, automask, kind, k_variant, seed)
    "mono_automask":   (2, 32, 64, [0, -1, 1], True, "smooth", "monodepth2", 0),
    "mono_iid":        (1, 32, 64, [0, -1, 1], True, "iid", "floor", 1),
    "stereo_automask": (1, 32, 64, [0, -1, 1, "s"], True, "smooth", "row1_width", 2),
    "mono_nomask":     (1, 32, 64, [0, -1, 1], False, "smooth", "monodepth2", 3),
    "single_nomask":   (1, 32, 64, [0, 1], False, "smooth", "monodepth2", 4),
    "five_frames":     (1, 32, 64, [0, -2, -1, 1, 2], True, "smooth", "monodepth2", 5),   # S = 4
    "partial_tiles":   (2, 40, 72, [0, -1, 1], True, "smooth", "floor", 6),   # H, W not multiples of the 32x16 tile
}


def run_reference(B, H, W, frame_ids, automask, kind, k_variant, seed):
    torch.manual_seed(seed)
    inputs, outputs = syn.make_batch(B, H, W, frame_ids, 4, seed, kind, k_variant)
    srcs = frame_ids[1:]
    Ts = {}
    for f in srcs:
        if f == "s":
            continue
        M = param2matrix(outputs[("axisangle", f)].detach(), outputs[("translation", f)].detach(),
                         invert=(f < 0)).detach().clone().requires_grad_(True)
        outputs[("c2c", f, 0)] = M
        Ts[f] = M
    opt = SimpleNamespace(frame_ids=frame_ids, scales=range(4), height=H, width=W, min_depth=0.1,
                          max_depth=100.0, pose_type="separate", pose_frames=2,
                          use_automasking=automask, disp_smoothness=1e-3, batch=B)
    setting = SimpleNamespace(inv_projection={0: Depth2PointCloud(B, H, W)},
                              for_projection={0: PointCloud2Pixel(B, H, W)},
                              loss={"reprojection": ReprojectionLoss(), "edge_aware": SmoothLoss()})
    c = compute(opt, "cpu")

    noise_rec, min_rec = [], []
    orig_randn, orig_min = torch.randn, torch.min

    def rec_randn(*a, **k):
        r = orig_randn(*a, **k)
        noise_rec.append(r.clone())
        return r

    def rec_min(*a, **k):
        r = orig_min(*a, **k)
        if isinstance(r, tuple) or hasattr(r, "indices"):
            min_rec.append((r[0].detach().clone(), r[1].detach().clone()))
        return r

    torch.randn, torch.min = rec_randn, rec_min
    try:
        c.image2warping(inputs, outputs, setting)
        c.compute_loss(inputs, outputs, setting)
    finally:
        torch.randn, torch.min = orig_randn, orig_min
    loss = outputs["loss"]
    loss.backward()

    d = {"meta_B": B, "meta_H": H, "meta_W": W, "meta_automask": int(automask),
         "meta_frame_ids": np.array([str(f) for f in frame_ids])}
    d["target"] = inputs[("color", 0, 0)].numpy()
    for i, f in enumerate(srcs):
        d[f"source{i}"] = inputs[("color", f, 0)].numpy()
        if f == "s":
            d[f"T{i}"] = inputs["stereo"].numpy()
        else:
            d[f"T{i}"] = Ts[f].detach().numpy()
            d[f"grad_T{i}"] = Ts[f].grad.numpy()
    for s in range(4):
        d[f"disp{s}"] = outputs[("disp", s)].detach().numpy()
        d[f"color_pyr{s}"] = inputs[("color", 0, s)].numpy()
        d[f"grad_disp{s}"] = outputs[("disp", s)].grad.numpy()
        d[f"depth{s}"] = outputs[("depth", 0, s)].detach().numpy()
        if automask:
            d[f"noise{s}"] = noise_rec[s].numpy()
        if len(min_rec) == 4:
            d[f"per_pixel{s}"] = min_rec[s][0].numpy()
            d[f"argmin{s}"] = min_rec[s][1].numpy().astype(np.uint8)
    d["K"] = inputs[("K", 0)].numpy()
    d["inv_K"] = inputs[("inv_K", 0)].numpy()
    d["loss"] = loss.detach().numpy()
    return d


def pose_case(seed):
    """Known-answer vectors for param2matrix (model_layer/warp.py:126-153)."""
    g = torch.Generator().manual_seed(seed)
    aa = (0.3 * torch.randn(6, 1, 3, generator=g)).requires_grad_(True)
    tr = (0.5 * torch.randn(6, 1, 3, generator=g)).requires_grad_(True)
    aa.data[0] = 0.0  # the |aa| -> 0 corner (axis = aa / (|aa| + 1e-5))
    cot = torch.randn(2, 6, 4, 4, generator=g)
    d = {"aa": aa.detach().numpy(), "tr": tr.detach().numpy(), "cot": cot.numpy()}
    for k, inv in enumerate([False, True]):
        M = param2matrix(aa, tr, invert=inv)
        ga, gt = torch.autograd.grad((M * cot[k]).sum(), [aa, tr])
        d[f"M{k}"] = M.detach().numpy()
        d[f"grad_aa{k}"] = ga.numpy()
        d[f"grad_tr{k}"] = gt.numpy()
    return d


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in CASES.items():
        d = run_reference(*cfg)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "loss", float(d["loss"]), os.path.getsize(path) // 1024, "KiB")
    np.savez_compressed(os.path.join(HERE, "pose.npz"), **pose_case(11))
    print("pose done")
