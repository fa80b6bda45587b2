"""Colour pyramid: pins the numpy restatement of Pillow's resampler against Pillow itself and
against vectors produced by the reference dataset's own resize / ToTensor objects (tests/golden/pyramid.npz),
and checks the library's HOST coefficient tables (md2_pyramid_tables_fill, no GPU) against the restatement."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from oracle import oracle_resize as R


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "pyramid.npz"))


def test_oracle_matches_reference_golden():
    z = _golden()
    img = z["image"]
    for flip in (0, 1):
        got = R.color_pyramid(img, 64, 192, 4, flip=bool(flip))
        for s in range(4):
            ref = z[f"color_f{flip}_s{s}"]
            assert got[s].shape == ref.shape and got[s].dtype == ref.dtype
            assert np.array_equal(got[s], ref), (flip, s, np.abs(got[s] - ref).max())


@pytest.mark.parametrize("Hin,Win,H,W", [(375, 1242, 192, 640), (375, 1242, 24, 80), (100, 333, 64, 96), (64, 96, 64, 96),
                                         (40, 50, 64, 96)])
def test_oracle_matches_pillow(Hin, Win, H, W):
    from PIL import Image
    img = np.random.default_rng(Hin + W).integers(0, 256, (Hin, Win, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((W, H), Image.LANCZOS))
    assert np.array_equal(R.resize_antialias(img, H, W), ref)


def test_host_tables_match_oracle():
    """Layout of csrc/md2_pipeline.cu: per level, X axis as first-aligned-word + three packed signed-digit words per
    aligned input word (c = d0 + 2^8 d1 + 2^16 d2), Y axis as Pillow's bounds + coefficients."""
    import md2_b200.build as b
    import md2_b200.pipeline as P
    b.build_cuda_library()
    for (Hin, Win, H, W, scales) in [(375, 1242, 192, 640, 4), (370, 1226, 320, 1024, 4), (120, 400, 64, 192, 4), (64, 96, 64, 96, 1)]:
        cfg, tab = P.pyramid_tables(3, Hin, Win, H, W, scales)
        off = 0
        for s in range(scales):
            wout, hout = W >> s, H >> s
            bounds, kk, ksize = R.precompute_coeffs(Win, wout)
            kw = (ksize + 3) // 4 + 1
            w0 = tab[off:off + wout]
            off += wout
            cd = tab[off:off + 3 * kw * wout].reshape(3, kw, wout)
            off += 3 * kw * wout
            assert np.array_equal(w0, bounds[:, 0] >> 2)
            digits = np.stack([((cd.view(np.uint32) >> (8 * bb)) & 255).astype(np.uint8).view(np.int8).astype(np.int64)
                               for bb in range(4)], -1)            # [3, kw, wout, 4]
            coef = digits[0] + 256 * digits[1] + 65536 * digits[2]   # [kw, wout, 4]
            assert np.abs(digits[2]).max() <= 64
            for x in range(wout):
                full = coef[:, x, :].reshape(-1)                     # bytes of the aligned words, from 4 * w0
                lead = bounds[x, 0] - 4 * w0[x]
                assert np.array_equal(full[lead:lead + bounds[x, 1]], kk[x, :bounds[x, 1]])
                assert not full[:lead].any() and not full[lead + bounds[x, 1]:].any()
            by, ky, ksy = R.precompute_coeffs(Hin, hout)
            assert np.array_equal(tab[off:off + 2 * hout].reshape(hout, 2), by)
            off += 2 * hout
            assert np.array_equal(tab[off:off + hout * ksy].reshape(hout, ksy), ky)
            off += hout * ksy
            ng = (ksy + 3) // 4                                      # the same taps, digit-packed four rows at a time
            ycd = tab[off:off + hout * ng * 3].reshape(hout, ng, 3)
            off += hout * ng * 3
            yd = np.stack([((ycd.view(np.uint32) >> (8 * bb)) & 255).astype(np.uint8).view(np.int8).astype(np.int64)
                           for bb in range(4)], -1)                  # [hout, ng, 3, 4]
            ycoef = (yd[:, :, 0] + 256 * yd[:, :, 1] + 65536 * yd[:, :, 2]).reshape(hout, ng * 4)
            for y in range(hout):
                assert np.array_equal(ycoef[y, :by[y, 1]], ky[y, :by[y, 1]]) and not ycoef[y, by[y, 1]:].any()
        assert off == tab.size


def test_pyramid_abi_validates_without_gpu():
    import md2_b200.cabi as cabi
    lib = cabi.load_library()
    assert lib.md2_pyramid_tables_bytes(C.byref(cabi.md2_pyramid_cfg(0, 375, 1242, 192, 640, 4))) == 0
    assert lib.md2_pyramid_tables_bytes(C.byref(cabi.md2_pyramid_cfg(1, 375, 1242, 192, 640, 5))) == 0
    assert lib.md2_pyramid_workspace_bytes(C.byref(cabi.md2_pyramid_cfg(2, 375, 1242, 192, 640, 4))) >= 2 * 3 * 375 * (1242 + 640)
    cfg = cabi.md2_pyramid_cfg(1, 375, 1242, 192, 640, 4)
    assert lib.md2_pyramid_tables_fill(C.byref(cfg), C.c_void_p(0)) == cabi.MD2_ERR_NULL
    one = C.c_void_p(8)
    outs = (C.c_void_p * 4)(8, 8, 8, 0)
    assert lib.md2_color_pyramid(C.byref(cfg), one, None, one, outs, one, None) == cabi.MD2_ERR_NULL
    assert lib.md2_color_pyramid(C.byref(cfg), one, None, one, outs, None, None) == cabi.MD2_ERR_WORKSPACE


def test_intrinsics_match_reference_golden():
    import md2_b200.pipeline as P
    z = _golden()
    intr = P.resize_intrinsic(192, 64, 4, "row1_width")
    for s in range(4):
        assert np.array_equal(intr[("K", s)].numpy(), z[f"K{s}"])
        assert np.array_equal(intr[("inv_K", s)].numpy(), z[f"inv_K{s}"])


def test_to_tensor_oracle_is_torchvisions_totensor():
    """kitti_mono.py:283 - transforms.ToTensor() on the loader's PIL images: every byte value in every channel, and
    the golden level-0 image the reference's own objects produced (tests/golden/pyramid.npz)."""
    import torch
    from PIL import Image
    from torchvision import transforms
    img = np.arange(256, dtype=np.uint8).repeat(3 * 5).reshape(16, 80, 3)[:, ::-1].copy()
    img[..., 1] = img[..., 1][::-1]
    ref = transforms.ToTensor()(Image.fromarray(img)).numpy()
    got = R.to_tensor(img)
    assert got.dtype == np.float32 and got.shape == (3, 16, 80) and np.array_equal(got, ref)
    batch = np.random.default_rng(3).integers(0, 256, (2, 7, 9, 3), dtype=np.uint8)
    ref_b = np.stack([transforms.ToTensor()(Image.fromarray(b)).numpy() for b in batch])
    assert np.array_equal(R.to_tensor(batch), ref_b)
    z = _golden()
    lvl0 = R.resize_antialias(z["image"], 64, 192)
    assert np.array_equal(R.to_tensor(lvl0), z["color_f0_s0"])


def test_to_tensor_entry_point_validates_without_gpu():
    """include/md2_pipeline.h md2_to_tensor: argument errors and empty work return before any launch."""
    import md2_b200.build as b
    import md2_b200.cabi as cabi
    b.build_cuda_library()
    lib = cabi.load_library()
    one = cabi.md2_u8_images(8, 8, 1, 4, 4)
    arr = (cabi.md2_u8_images * 1)(one)
    assert lib.md2_to_tensor(0, None, None) == 0
    assert lib.md2_to_tensor(1, None, None) == cabi.MD2_ERR_NULL
    assert lib.md2_to_tensor(cabi.MD2_TO_TENSOR_MAX + 1, arr, None) == cabi.MD2_ERR_SHAPE
    assert lib.md2_to_tensor(-1, arr, None) == cabi.MD2_ERR_SHAPE
    bad = (cabi.md2_u8_images * 1)(cabi.md2_u8_images(8, 8, 1, -4, 4))
    assert lib.md2_to_tensor(1, bad, None) == cabi.MD2_ERR_SHAPE
    null_src = (cabi.md2_u8_images * 1)(cabi.md2_u8_images(0, 8, 1, 4, 4))
    assert lib.md2_to_tensor(1, null_src, None) == cabi.MD2_ERR_NULL
    empty = (cabi.md2_u8_images * 2)(cabi.md2_u8_images(0, 0, 0, 4, 4), cabi.md2_u8_images(0, 0, 3, 0, 4))
    assert lib.md2_to_tensor(2, empty, None) == 0            # nothing to convert, nothing launched
