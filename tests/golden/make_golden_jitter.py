"""Golden vectors for the colour-jitter branch (model_loader/kitti_mono.py:281-282, 351-357).

Run in the build container:   python tests/golden/make_golden_jitter.py
The reference calls transforms.ColorJitter.get_params(...) once and then CALLS the result on PIL images - the
behaviour of torchvision <= 0.8, where get_params returned a Compose of the four adjust_* Lambdas in shuffled
order.  torchvision 0.26 (this image) returns a tuple instead, so the reference line itself cannot run here; the
vectors are produced with what that Compose did: torchvision.transforms.functional.adjust_* applied to the PIL
image (functional_pil / Pillow), in a fixed shuffled order, on the same synthetic frame as pyramid.npz's level 0.
"""
import os

import numpy as np
import torchvision.transforms.functional as F
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [((2, 0, 3, 1), 1.13, 0.87, 1.19, -0.07), ((0, 1, 2, 3), 0.8, 1.2, 0.8, 0.1), ((3, 2, 1, 0), 1.0, 1.0, 1.0, 0.0),
         ((1, 3, 0, 2), 0.95, 1.05, 1.11, 0.033)]


def main():
    z = np.load(os.path.join(HERE, "pyramid.npz"))
    img = np.round(z["color_f0_s1"].transpose(1, 2, 0) * 255).astype(np.uint8)      # 32 x 96 RGB
    img[0, :6] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128]]
    fn = (F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue)
    d = {"image": img, "params": np.array([[*o, b, c, s, h] for o, b, c, s, h in CASES], dtype=np.float64)}
    for i, (order, b, c, s, h) in enumerate(CASES):
        pil = Image.fromarray(img)
        for k in order:
            pil = fn[k](pil, (b, c, s, h)[k])
        d[f"out{i}"] = np.array(pil)
    path = os.path.join(HERE, "jitter.npz")
    np.savez_compressed(path, **d)
    print("jitter.npz", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
