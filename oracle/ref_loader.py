"""Import the staged, unmodified reference (oracle/_ref, see stage_ref.py) - test / baseline infrastructure.

The reference imports plotting / augmentation packages at module top that this image lacks and the loss path never
uses; they are stubbed in sys.modules exactly as tests/golden/make_golden.py does.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_loaded = None


def available():
    return os.path.isfile(os.path.join(REF, "model_tool", "processor.py"))


def load():
    """-> namespace(compute, Depth2PointCloud, PointCloud2Pixel, param2matrix, ReprojectionLoss, SmoothLoss)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run python oracle/stage_ref.py where /root/reference exists)")
    sys.dont_write_bytecode = True
    for m in ["matplotlib", "matplotlib.pyplot", "albumentations", "albumentations.pytorch",
              "albumentations.pytorch.transforms", "albumentations.augmentations",
              "albumentations.augmentations.transforms", "skimage", "skimage.transform"]:
        sys.modules.setdefault(m, MagicMock())
    sys.modules["albumentations"].__version__ = "0.5.2"
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from model_tool.processor import compute
    from model_layer import Depth2PointCloud, PointCloud2Pixel, param2matrix
    from model_loss import ReprojectionLoss, SmoothLoss
    _loaded = SimpleNamespace(compute=compute, Depth2PointCloud=Depth2PointCloud, PointCloud2Pixel=PointCloud2Pixel,
                              param2matrix=param2matrix, ReprojectionLoss=ReprojectionLoss, SmoothLoss=SmoothLoss)
    return _loaded


def make_step(B, H, W, frame_ids, device, num_scales=4, seed=0, kind="iid", host_noise=True):
    """One training-loss step of the reference itself - compute.image2warping + compute.compute_loss
    (model_tool/processor.py:139-218) + loss.backward() (model_train.py:68) - on `device`, on a synthetic batch.

    host_noise=True is the reference as shipped: the auto-mask noise is a host torch.randn copied to the device
    every scale (processor.py:195).  host_noise=False times the same code with that draw made on the device
    (torch.randn patched to allocate there), i.e. without the host RNG and the H2D copy."""
    import torch
    import md2_b200.synthetic as syn
    R = load()
    inputs, outputs = syn.make_batch(B, H, W, frame_ids, num_scales, seed, kind, requires_grad=False)
    dev = torch.device(device)
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    aa = {f: outputs[("axisangle", f)].to(dev) for f in frame_ids[1:] if f != "s"}
    tr = {f: outputs[("translation", f)].to(dev) for f in frame_ids[1:] if f != "s"}
    Ts = {f: R.param2matrix(aa[f], tr[f], invert=(f < 0)).detach() for f in aa}
    disp0 = [outputs[("disp", s)].to(dev) for s in range(num_scales)]
    opt = SimpleNamespace(frame_ids=frame_ids, scales=range(num_scales), height=H, width=W, min_depth=0.1,
                          max_depth=100.0, pose_type="separate", pose_frames=2, use_automasking=True,
                          disp_smoothness=1e-3, batch=B)
    setting = SimpleNamespace(inv_projection={0: R.Depth2PointCloud(B, H, W).to(dev)},
                              for_projection={0: R.PointCloud2Pixel(B, H, W).to(dev)},
                              loss={"reprojection": R.ReprojectionLoss().to(dev), "edge_aware": R.SmoothLoss().to(dev)})
    comp = R.compute(opt, dev)
    orig_randn = torch.randn

    def device_randn(*a, **k):
        k.setdefault("device", dev)
        return orig_randn(*a, **k)

    def step():
        outs = {("disp", s): disp0[s].clone().requires_grad_(True) for s in range(num_scales)}
        for f, T in Ts.items():
            outs[("c2c", f, 0)] = T.clone().requires_grad_(True)
        if not host_noise:
            torch.randn = device_randn
        try:
            comp.image2warping(inputs, outs, setting)
            comp.compute_loss(inputs, outs, setting)
        finally:
            torch.randn = orig_randn
        outs["loss"].backward()
        return outs["loss"].detach()
    return step
