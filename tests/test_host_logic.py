"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol that
include/md2_loss.h declares, argument validation, workspace sizing, and the package refuses to
run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

import md2_b200
import md2_b200.cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import md2_b200.build as b
    b.build_cuda_library()
    return cabi.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))))
    declared = set(re.findall(r"\b(md2_[a-z_0-9]+)\s*\(", header))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.md2_version()


def test_operator_entry_points_validate_without_gpu(lib):
    """include/md2_ops.h: argument errors are reported before anything is launched."""
    null = C.c_void_p(0)
    one = C.c_void_p(8)  # never dereferenced: the shape check comes first
    assert lib.md2_disp2depth_forward(0, one, 0.1, 100.0, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_disp2depth_forward(4, one, 0.0, 100.0, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_disp2depth_forward(4, null, 0.1, 100.0, one, one, null) == cabi.MD2_ERR_NULL
    assert lib.md2_upsample_forward(1, 0, 4, 8, 8, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_upsample_backward(1, 4, 4, 8, 8, null, one, null) == cabi.MD2_ERR_NULL
    assert lib.md2_backproject_forward(0, 8, 8, one, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_project_forward(1, 1, 8, one, one, one, 1e-7, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_project_backward(1, 8, 8, one, one, one, 1e-7, one, one, null, null) == cabi.MD2_ERR_NULL
    assert lib.md2_grid_sample_forward(1, 0, 8, 8, 8, 8, one, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_reflection_pad2d_forward(1, 4, 4, 4, 4, 0, 0, 0, 1, one, one, null) == cabi.MD2_ERR_SHAPE  # pad >= W
    assert lib.md2_reflection_pad2d_forward(1, 4, 4, 4, 1, 1, 1, -1, 0, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_reflection_pad2d_forward(1, 4, 4, 4, 1, 1, 1, 1, 1, null, one, null) == cabi.MD2_ERR_NULL
    assert lib.md2_reflection_pad2d_backward(0, 4, 4, 4, 1, 1, 1, 1, 1, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_reflection_pad2d_backward(1, 4, 4, 4, 1, 1, 1, 1, 0, one, null, null) == cabi.MD2_ERR_NULL
    assert lib.md2_maxpool2d_nhwc_forward(1, 4, 4, 4, 3, 2, 2, one, one, one, null) == cabi.MD2_ERR_SHAPE     # 2p > k
    assert lib.md2_maxpool2d_nhwc_forward(1, 4, 2, 2, 5, 1, 0, one, one, one, null) == cabi.MD2_ERR_SHAPE     # window > image
    assert lib.md2_maxpool2d_nhwc_forward(1, 4, 8, 8, 3, 2, 1, one, one, null, null) == cabi.MD2_ERR_NULL
    assert lib.md2_maxpool2d_nhwc_backward(1, 4, 8, 8, 3, 0, 1, one, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_maxpool2d_nhwc_backward(1, 4, 8, 8, 3, 2, 1, null, one, one, null) == cabi.MD2_ERR_NULL
    assert lib.md2_reprojection_forward(1, 2, 8, one, one, one, null) == cabi.MD2_ERR_SHAPE
    assert lib.md2_smooth_forward(1, 8, 8, one, one, null, one, null) == cabi.MD2_ERR_NULL


def test_workspace_and_validation_without_gpu(lib):
    ok = cabi.make_cfg(12, 192, 640, 2)
    assert lib.md2_workspace_bytes(C.byref(ok)) > 0
    assert lib.md2_launches_per_step(C.byref(ok), 1) == 5
    assert lib.md2_launches_per_step(C.byref(ok), 0) == 3
    for bad in (cabi.make_cfg(0, 192, 640, 2), cabi.make_cfg(12, 190, 640, 2), cabi.make_cfg(12, 192, 640, 5),
                cabi.make_cfg(12, 192, 640, 2, num_scales=5), cabi.make_cfg(12, 192, 640, 2, min_depth=0.0)):
        assert lib.md2_workspace_bytes(C.byref(bad)) == 0
    # NULL structs are rejected before any launch
    assert lib.md2_loss_forward(C.byref(ok), None, None, None, None) < 0


def test_struct_layout_matches_header():
    # 6 ints + 4 doubles; pointer arrays of MD2_MAX_SOURCES / MD2_MAX_SCALES
    assert C.sizeof(cabi.md2_cfg) == 6 * 4 + 4 * 8
    assert C.sizeof(cabi.md2_inputs) == 8 * (1 + 4 + 4 + 4 + 2 + 4 + 4 + 1 + 1)  # ... seed, seed_dev
    assert C.sizeof(cabi.md2_outputs) == 32
    assert C.sizeof(cabi.md2_grads) == 64


def test_no_cpu_fallback():
    from md2_b200 import functional as F_
    t = torch.zeros(1, 3, 32, 64)
    with pytest.raises(RuntimeError):
        F_.view_synthesis_loss(t, [t], [torch.zeros(1, 1, 32, 64)], [t], torch.eye(4)[None], torch.eye(4)[None],
                               [torch.eye(4)[None]])


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ or the host emulation."""
    pkg = os.path.dirname(md2_b200.__file__)
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle_torch|#include[^\n]*oracle|import[^\n]*host_emu|libmd2emu)")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), f


def test_channels_last_module_swaps_are_host_logic():
    """md2_b200.modules.use_channels_last_padding / _pooling only rewire modules (no GPU needed); the swapped modules
    refuse host tensors - there is no CPU path behind them."""
    import torch
    import torch.nn as nn
    import md2_b200.modules as M

    class Conv3x3(nn.Module):                       # shaped like model_layer/depth_decoder.py:36-50
        def __init__(self):
            super().__init__()
            self.pad = nn.ReflectionPad2d(1)
            self.conv = nn.Conv2d(4, 4, 3)

    net = nn.ModuleDict({"a": Conv3x3(), "b": nn.Sequential(Conv3x3(), nn.ZeroPad2d(1)),
                         "pool": nn.MaxPool2d(3, 2, 1), "pool_ceil": nn.MaxPool2d(3, 2, 1, ceil_mode=True),
                         "pool_rect": nn.MaxPool2d((3, 2), 2), "pool_dil": nn.MaxPool2d(3, 2, 1, dilation=2)})
    assert M.use_channels_last_padding(net) == 2
    assert isinstance(net["a"].pad, M.ReflectionPad2d) and net["a"].pad.padding == (1, 1, 1, 1)
    assert isinstance(net["b"][1], nn.ZeroPad2d)                     # only reflection pads are replaced
    assert M.use_channels_last_padding(net) == 0                     # idempotent
    assert M.use_channels_last_pooling(net) == 1
    assert isinstance(net["pool"], M.MaxPool2d) and (net["pool"].kernel_size, net["pool"].stride, net["pool"].padding) == (3, 2, 1)
    for k in ("pool_ceil", "pool_rect", "pool_dil"):                 # unsupported variants keep ATen's operator
        assert type(net[k]) is nn.MaxPool2d
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net["a"].pad(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net["pool"](x)
    with pytest.raises(ValueError):
        M.ReflectionPad2d((1, 2, 3))
    with pytest.raises(ValueError):
        M.MaxPool2d(3, 2, 2)
