"""Golden vectors for the colour pyramid, produced with the reference dataset's own objects.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_pyramid.py
KITTIMonoDataset_v2 (model_loader/kitti_mono.py:255-288) is instantiated unmodified; the only shim is
Image.ANTIALIAS = Image.LANCZOS, the alias Pillow removed in 10.0 (the reference predates that).
Its self.resize[scale] and self.numpy2tensor are applied to a synthetic "decoded frame" exactly as
__getitem__ does (:352-355), with and without the left-right flip of load_image (:303-304).
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MD2_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from PIL import Image  # noqa: E402

if not hasattr(Image, "ANTIALIAS"):
    Image.ANTIALIAS = Image.LANCZOS
for _m in ["matplotlib", "matplotlib.pyplot", "albumentations", "albumentations.pytorch", "albumentations.pytorch.transforms",
           "albumentations.augmentations", "albumentations.augmentations.transforms", "skimage", "skimage.transform", "cv2"]:
    sys.modules[_m] = MagicMock()
sys.modules["albumentations"].__version__ = "0.5.2"

from model_loader.kitti_mono import KITTIMonoDataset_v2  # noqa: E402

Hin, Win, H, W = 120, 400, 64, 192


def main():
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:Hin, 0:Win]
    base = 128 + 90 * np.sin(xx / 17.0)[..., None] * np.cos(yy / 11.0)[..., None] * np.array([1.0, 0.7, -0.8])
    img = np.clip(base + rng.normal(0, 25, (Hin, Win, 3)), 0, 255).astype(np.uint8)
    img[:8, :8] = 255      # saturated corners: clipping of the Lanczos overshoot
    img[-8:, -8:] = 0
    ds = KITTIMonoDataset_v2("", [], False, [0], height=H, width=W, scale=4)
    d = {"image": img}
    for flip in (0, 1):
        pil = Image.fromarray(img)
        if flip:
            pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
        for s in range(4):
            d[f"color_f{flip}_s{s}"] = ds.numpy2tensor(ds.resize[s](pil)).numpy()
    intr = ds.resize_intrinsic({})
    for s in range(4):
        d[f"K{s}"], d[f"inv_K{s}"] = intr[("K", s)].numpy(), intr[("inv_K", s)].numpy()
    path = os.path.join(HERE, "pyramid.npz")
    np.savez_compressed(path, **d)
    print("pyramid.npz", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
