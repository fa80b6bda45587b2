"""Pins oracle/oracle_jitter.py against the installed Pillow / torchvision (every adjustment on random images,
both HSV conversions exhaustively over all 2^24 colours) and against tests/golden/jitter.npz."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from oracle import oracle_jitter as J


def test_against_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "jitter.npz"))
    for i, row in enumerate(z["params"]):
        order = [int(v) for v in row[:4]]
        got = J.color_jitter(z["image"], order, *row[4:])
        assert np.array_equal(got, z[f"out{i}"]), i


@pytest.mark.parametrize("factor", [0.8, 0.9137, 1.0, 1.0421, 1.2])
def test_adjustments_match_torchvision_pil(factor):
    import torchvision.transforms.functional as F
    from PIL import Image
    img = np.random.default_rng(int(factor * 1000)).integers(0, 256, (48, 80, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    assert np.array_equal(J.adjust_brightness(img, factor), np.array(F.adjust_brightness(pil, factor)))
    assert np.array_equal(J.adjust_contrast(img, factor), np.array(F.adjust_contrast(pil, factor)))
    assert np.array_equal(J.adjust_saturation(img, factor), np.array(F.adjust_saturation(pil, factor)))
    hue = (factor - 1.0) / 2.0
    assert np.array_equal(J.adjust_hue(img, hue), np.array(F.adjust_hue(pil, hue)))


def test_hsv_conversions_exhaustive():
    from PIL import Image
    a = np.arange(256, dtype=np.uint8)
    c = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(4096, 4096, 3)
    assert np.array_equal(J.rgb_to_hsv(c), np.array(Image.fromarray(c).convert("HSV")))
    assert np.array_equal(J.hsv_to_rgb(c), np.array(Image.fromarray(c, "HSV").convert("RGB")))


def test_jitter_abi_validates_without_gpu():
    import ctypes as C
    import md2_b200.build as b
    import md2_b200.cabi as cabi
    b.build_cuda_library()
    lib = cabi.load_library()
    mk = lambda N=1, H=8, W=8, order=(0, 1, 2, 3), f=(1.0, 1.0, 1.0, 0.0): cabi.md2_jitter_cfg(N, H, W, (C.c_int * 4)(*order), *f)
    assert lib.md2_jitter_workspace_bytes(C.byref(mk())) >= 3 * 64 + 8
    assert lib.md2_jitter_workspace_bytes(C.byref(mk(N=0))) == 0
    assert lib.md2_jitter_workspace_bytes(C.byref(mk(order=(0, 1, 2, 4)))) == 0
    assert lib.md2_jitter_workspace_bytes(C.byref(mk(f=(1.0, 1.0, 1.0, 0.6)))) == 0
    assert lib.md2_jitter_workspace_bytes(C.byref(mk(f=(-0.1, 1.0, 1.0, 0.0)))) == 0
    one = C.c_void_p(8)
    assert lib.md2_color_jitter(C.byref(mk()), None, None, one, one, None) == cabi.MD2_ERR_NULL
    assert lib.md2_color_jitter(C.byref(mk()), one, None, one, None, None) == cabi.MD2_ERR_WORKSPACE
    assert lib.md2_color_jitter(C.byref(mk(order=(9, 0, 0, 0))), one, None, one, one, None) == cabi.MD2_ERR_CONFIG
