"""GPU parity of the colour-pyramid kernels (md2_b200.pipeline -> include/md2_pipeline.h -> csrc/md2_pipeline.cu):
bit-exact (integer work) against the reference dataset's own outputs (tests/golden/pyramid.npz), against the
numpy oracle and, at the full KITTI size, against Pillow itself."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import oracle_resize as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pyramid_matches_reference_golden():
    import md2_b200.pipeline as P
    z = np.load(os.path.join(GOLDEN_DIR, "pyramid.npz"))
    img = torch.from_numpy(z["image"]).to(DEV)
    pyr = P.ColorPyramid(2, 120, 400, 64, 192, 4, device=DEV)
    levels = pyr(torch.stack([img, img]), flip=torch.tensor([0, 1]))
    for s in range(4):
        assert levels[s].shape == (2, 3, 64 >> s, 192 >> s) and levels[s].dtype == torch.float32
        for flip in (0, 1):
            assert np.array_equal(levels[s][flip].cpu().numpy(), z[f"color_f{flip}_s{s}"]), (s, flip)


@pytest.mark.parametrize("N,Hin,Win,H,W", [(3, 375, 1242, 192, 640), (2, 370, 1226, 320, 1024), (1, 64, 96, 64, 96)])
def test_pyramid_matches_pillow_at_kitti_size(N, Hin, Win, H, W):
    from PIL import Image
    import md2_b200.pipeline as P
    rng = np.random.default_rng(N + W)
    imgs = rng.integers(0, 256, (N, Hin, Win, 3), dtype=np.uint8)
    flip = np.array([i % 2 for i in range(N)], dtype=np.uint8)
    levels = P.ColorPyramid(N, Hin, Win, H, W, 4, device=DEV)(torch.from_numpy(imgs).to(DEV), torch.from_numpy(flip))
    for n in range(N):
        pil = Image.fromarray(imgs[n])
        if flip[n]:
            pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
        for s in range(4):
            ref = np.array(pil.resize((W >> s, H >> s), Image.LANCZOS)).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
            assert np.array_equal(levels[s][n].cpu().numpy(), ref), (n, s)


def test_pyramid_matches_oracle_and_rejects_bad_input():
    import md2_b200.pipeline as P
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (1, 100, 333, 3), dtype=np.uint8)
    pyr = P.ColorPyramid(1, 100, 333, 64, 96, 3, device=DEV)
    levels = pyr(torch.from_numpy(img).to(DEV))
    ref = R.color_pyramid(img[0], 64, 96, 3)
    for s in range(3):
        assert np.array_equal(levels[s][0].cpu().numpy(), ref[s])
    # a level whose width is not a multiple of 4 takes the scalar vertical pass (200, 100, 50)
    img2 = rng.integers(0, 256, (2, 90, 301, 3), dtype=np.uint8)
    lv2 = P.ColorPyramid(2, 90, 301, 40, 200, 3, device=DEV)(torch.from_numpy(img2).to(DEV), torch.tensor([1, 0]))
    for n, fl in ((0, True), (1, False)):
        ref2 = R.color_pyramid(img2[n], 40, 200, 3, flip=fl)
        for s in range(3):
            assert np.array_equal(lv2[s][n].cpu().numpy(), ref2[s]), (n, s)
    with pytest.raises(RuntimeError):
        pyr(torch.from_numpy(img))                       # host tensor
    with pytest.raises(RuntimeError):
        pyr(torch.from_numpy(img).to(DEV).float())       # wrong dtype
    with pytest.raises(RuntimeError):
        P.ColorPyramid(1, 100, 333, 64, 96, 3, device="cpu")


def test_every_byte_value_converts_like_totensor():
    """uint8 -> float32 / 255 for all 256 values (an identity-size 'resize' is the identity on the bytes)."""
    import md2_b200.pipeline as P
    img = np.arange(256, dtype=np.uint8).repeat(3 * 4).reshape(1, 16, 64, 3)   # every value, all channels
    out = P.ColorPyramid(1, 16, 64, 16, 64, 1, device=DEV)(torch.from_numpy(img).to(DEV))[0][0].cpu().numpy()
    ref = img[0].transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    assert np.array_equal(out, ref)


def test_to_tensor_matches_oracle_bit_exact():
    """md2_to_tensor (kitti_mono.py:283 on the device): every byte value; row lengths that are / are not multiples
    of four pixels; several groups in one launch, more groups than one launch takes, an empty group."""
    import md2_b200.pipeline as P
    from oracle import oracle_resize as R
    rng = np.random.default_rng(11)
    every = np.arange(256, dtype=np.uint8).repeat(3 * 4).reshape(1, 16, 64, 3)
    shapes = [(12, 192, 640), (12, 96, 320), (12, 48, 160), (12, 24, 80), (3, 7, 9), (1, 1, 1), (2, 5, 6), (0, 4, 4),
              (1, 33, 130)]
    imgs = [every] + [rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8) for n, h, w in shapes]
    dev = [torch.from_numpy(a).to(DEV) for a in imgs]
    outs = P.to_tensor(dev)
    assert len(outs) == len(imgs)
    for a, o in zip(imgs, outs):
        ref = R.to_tensor(a) if a.size else np.zeros((a.shape[0], 3, a.shape[1], a.shape[2]), np.float32)
        assert o.dtype == torch.float32 and tuple(o.shape) == ref.shape
        assert np.array_equal(o.cpu().numpy(), ref)
    # a single tensor in, a single tensor out; preallocated outputs; more than MD2_TO_TENSOR_MAX groups
    one = P.to_tensor(dev[1])
    assert torch.equal(one, outs[1])
    pre = torch.full_like(outs[2], -1.0)
    assert P.to_tensor(dev[2], out=pre) is pre and torch.equal(pre, outs[2])
    many = P.to_tensor([dev[5]] * 19 + [dev[6]])
    assert len(many) == 20 and all(torch.equal(m, outs[5]) for m in many[:19]) and torch.equal(many[19], outs[6])
    # an unaligned view of the bytes takes the scalar path and still matches
    flat = torch.zeros(3 * 8 * 12 * 3 + 1, dtype=torch.uint8, device=DEV)
    flat[1:] = torch.from_numpy(imgs[1].reshape(-1)[:3 * 8 * 12 * 3].copy()).to(DEV)
    shifted = flat[1:].view(3, 8, 12, 3)
    assert np.array_equal(P.to_tensor(shifted).cpu().numpy(), R.to_tensor(shifted.cpu().numpy()))
    with pytest.raises(RuntimeError):
        P.to_tensor(torch.from_numpy(imgs[1]))                    # host tensor: no CPU path
    with pytest.raises(RuntimeError):
        P.to_tensor(dev[1].float())                               # wrong dtype
    with pytest.raises(RuntimeError):
        P.to_tensor(dev[1].permute(0, 3, 1, 2))                   # not HWC


def test_loss_from_uploaded_bytes_equals_loss_from_float_tensors():
    """The training step that uploads bytes and converts on the device computes the same loss, per-pixel maps and
    argmin, bit for bit, as the one that receives the loader's float tensors (ToTensor on the host)."""
    import md2_b200.cabi as cabi
    import md2_b200.pipeline as P
    from oracle import oracle_resize as R
    from test_gpu_parity import synth_args
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W = 2, 64, 96
    args = synth_args(B, H, W, [0, -1, 1], True, "smooth", 21)
    rng = np.random.default_rng(5)
    u8 = {"pyr": [rng.integers(0, 256, (B, H >> s, W >> s, 3), dtype=np.uint8) for s in range(4)],
          "src": [rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8) for _ in range(2)]}
    host = dict(args)
    host["color_pyr"] = [torch.from_numpy(R.to_tensor(a)).to(DEV) for a in u8["pyr"]]
    host["target"] = host["color_pyr"][0]
    host["sources"] = [torch.from_numpy(R.to_tensor(a)).to(DEV) for a in u8["src"]]
    conv = P.to_tensor([torch.from_numpy(a).to(DEV) for a in u8["pyr"] + u8["src"]])
    devc = dict(args)
    devc["color_pyr"], devc["target"], devc["sources"] = conv[:4], conv[0], conv[4:]
    cl = cabi.CLoss()
    a, b = cl.forward_backward(host), cl.forward_backward(devc)
    assert float(a["loss"]) == float(b["loss"])
    assert torch.equal(a["per_pixel"], b["per_pixel"]) and torch.equal(a["argmin"], b["argmin"])
    from helpers import norm_rel
    for s in range(4):   # identical inputs; the gradient scatter uses float atomics, whose order is not fixed
        assert norm_rel(a["grad_disp"][s], b["grad_disp"][s]) <= 1e-6


def _jitter_ref(img_u8, order, b, c, s, h):
    import torchvision.transforms.functional as F
    from PIL import Image
    fn = (F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue)
    pil = Image.fromarray(img_u8)
    for k in order:
        pil = fn[k](pil, (b, c, s, h)[k])
    return np.array(pil)


def test_color_jitter_matches_golden_and_pil():
    """do_color branch (kitti_mono.py:351-357): bit-identical to torchvision's PIL adjustments, uint8 between steps."""
    import md2_b200.pipeline as P
    z = np.load(os.path.join(GOLDEN_DIR, "jitter.npz"))
    x = torch.from_numpy(z["image"].transpose(2, 0, 1).astype(np.float32) / np.float32(255))[None].to(DEV)
    for i, row in enumerate(z["params"]):
        out = P.ColorJitter([int(v) for v in row[:4]], *row[4:])(x)
        got = np.round(out[0].cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
        assert np.array_equal(got, z[f"out{i}"]), i
        assert np.array_equal(out[0].cpu().numpy(), z[f"out{i}"].transpose(2, 0, 1).astype(np.float32) / np.float32(255))
    # random parameters at the training size, two images with different content, one of them switched off
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (2, 192, 640, 3), dtype=np.uint8)
    x = torch.from_numpy(imgs.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)).to(DEV)
    import random
    random.seed(3)
    for _ in range(6):
        jit = P.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
        out = jit(x, apply=torch.tensor([1, 0]))
        got = np.round(out.cpu().numpy().transpose(0, 2, 3, 1) * 255).astype(np.uint8)
        assert np.array_equal(got[0], _jitter_ref(imgs[0], jit.order, *jit.factors)), (jit.order, jit.factors)
        assert np.array_equal(out[1].cpu().numpy(), x[1].cpu().numpy())


def test_color_jitter_hsv_exhaustive():
    """Every RGB colour through the hue adjustment (RGB -> HSV -> shift -> RGB) against Pillow."""
    import md2_b200.pipeline as P
    a = np.arange(256, dtype=np.uint8)
    c = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(4096, 4096, 3)
    x = torch.from_numpy(c.transpose(2, 0, 1).astype(np.float32) / np.float32(255))[None].to(DEV)
    for hue in (0.0, 0.0837, -0.1):
        out = P.ColorJitter([3], 1.0, 1.0, 1.0, hue)(x)
        got = torch.round(out[0] * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, _jitter_ref(c, [3], 1.0, 1.0, 1.0, hue)), hue
