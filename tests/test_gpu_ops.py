"""GPU parity of the symbol-level operators (md2_b200.modules -> include/md2_ops.h -> csrc/md2_l1.cu).

Forward: bit-exact against the oracle's per-operator functions (pinned against the reference's symbols by
tests/test_oracle_ops_golden.py) evaluated by ATen on the same GPU, and within fp32 noise of the reference's
recorded CPU outputs (tests/golden/ops.npz).  Backward: against autograd through the oracle, relative
tolerance 1e-4 on the gradient norm (north_star's gradient tolerance)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import oracle_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _golden():
    z = np.load(os.path.join(GOLDEN_DIR, "ops.npz"))
    return {k: torch.from_numpy(z[k]).to(DEV) for k in z.files}


def _close(a, b, tol):
    err = float((a.detach() - b.detach()).norm()) / max(float(b.detach().norm()), 1e-20)
    assert err <= tol, err


def _exact(a, b):
    a, b = a.detach(), b.detach()
    assert torch.equal(a, b), (float((a - b).abs().max()), int((a != b).sum()), a.numel())


@pytest.fixture(scope="module")
def M():
    import md2_b200.modules as m
    return m


def _rand(gen, *shape):
    return torch.rand(*shape, generator=gen, device=DEV)


@pytest.mark.parametrize("B,h,w,H,W", [(2, 12, 16, 24, 32), (3, 24, 80, 192, 640), (1, 96, 320, 192, 640), (2, 24, 32, 24, 32)])
def test_interpolate(M, B, h, w, H, W):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = _rand(g, B, 1, h, w).requires_grad_(True)
    y = M.interpolate(x, H, W, "bilinear", False)
    x2 = x.detach().clone().requires_grad_(True)
    ref = O.upsample_disp(x2, H, W)
    _exact(y, ref)
    cot = torch.randn(ref.shape, generator=g, device=DEV)
    _close(torch.autograd.grad((y * cot).sum(), x)[0], torch.autograd.grad((ref * cot).sum(), x2)[0], 1e-5)


def test_interpolate_rejects_other_modes(M):
    x = torch.zeros(1, 1, 4, 4, device=DEV)
    with pytest.raises(NotImplementedError):
        M.interpolate(x, 8, 8, "nearest", False)
    with pytest.raises(RuntimeError):
        M.interpolate(x.cpu(), 8, 8, "bilinear", False)
    with pytest.raises(RuntimeError):
        M.interpolate(x.double(), 8, 8, "bilinear", False)


def test_disparity2depth(M):
    g = torch.Generator(device=DEV).manual_seed(1)
    d = _rand(g, 3, 1, 48, 64).requires_grad_(True)
    d.data[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1e-6, 0.5], device=DEV)
    s, z = M.disparity2depth(d, 0.1, 100.0)
    d2 = d.detach().clone().requires_grad_(True)
    rs, rz = O.disp_to_depth(d2, 0.1, 100.0)
    _exact(s, rs)
    _exact(z, rz)
    c0, c1 = torch.randn_like(rs), torch.randn_like(rz)
    _close(torch.autograd.grad((s * c0).sum() + (z * c1).sum(), d)[0],
           torch.autograd.grad((rs * c0).sum() + (rz * c1).sum(), d2)[0], 1e-5)
    # only one of the two outputs used (processor.py:147 drops the scaled disparity)
    s, z = M.disparity2depth(d, 0.1, 100.0)
    _close(torch.autograd.grad(z.sum(), d)[0], torch.autograd.grad(O.disp_to_depth(d2, 0.1, 100.0)[1].sum(), d2)[0], 1e-5)


def _geometry(B, H, W, seed):
    import md2_b200.synthetic as syn
    g = torch.Generator(device=DEV).manual_seed(seed)
    K, inv_K = (t.to(DEV) for t in syn.make_intrinsics(B, H, W, "monodepth2"))
    depth = 0.5 + 8 * _rand(g, B, 1, H, W)
    T = torch.eye(4, device=DEV)[None].repeat(B, 1, 1)
    T[:, :3, :3] += 0.02 * torch.randn(B, 3, 3, generator=g, device=DEV)
    T[:, :3, 3] = 0.2 * torch.randn(B, 3, generator=g, device=DEV)
    return g, K, inv_K, depth, T


@pytest.mark.parametrize("B,H,W", [(2, 24, 32), (12, 192, 640), (1, 192, 640), (1, 320, 1024)])
def test_backproject_project(M, B, H, W):
    g, K, inv_K, depth, T = _geometry(B, H, W, 2)
    d1, d2 = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    T1, T2 = T.clone().requires_grad_(True), T.clone().requires_grad_(True)
    cam = M.Depth2PointCloud(B, H, W)(d1, inv_K)
    ref_cam = O.backproject(d2, inv_K, O.pixel_rays_grid(B, H, W, torch.float32, DEV))
    _exact(cam, ref_cam)
    grid = M.PointCloud2Pixel(B, H, W)(cam, K, T1)
    ref_grid = O.project(ref_cam, K, T2, H, W)
    _exact(grid, ref_grid)
    cot = torch.randn(ref_grid.shape, generator=g, device=DEV)
    gd, gT = torch.autograd.grad((grid * cot).sum(), [d1, T1])
    rd, rT = torch.autograd.grad((ref_grid * cot).sum(), [d2, T2])
    _close(gd, rd, 1e-4)
    _close(gT, rT, 1e-3)   # a sum of B*H*W fp32 terms with cancellation; autograd's own order is not better


@pytest.mark.parametrize("B,H,W", [(2, 24, 32), (4, 192, 640)])
def test_grid_sample(M, B, H, W):
    g = torch.Generator(device=DEV).manual_seed(3)
    img = _rand(g, B, 3, H, W)
    grid = (2.4 * _rand(g, B, H, W, 2) - 1.2)           # a tenth of the samples outside the image
    grid[0, 0, :4, 0] = torch.tensor([-1.0, 1.0, 0.0, 5.0], device=DEV)
    g1, g2 = grid.clone().requires_grad_(True), grid.clone().requires_grad_(True)
    out = M.grid_sample(img, g1, "border", True)
    ref = O.sample_border(img, g2)
    _exact(out, ref)
    cot = torch.randn(ref.shape, generator=g, device=DEV)
    _close(torch.autograd.grad((out * cot).sum(), g1)[0], torch.autograd.grad((ref * cot).sum(), g2)[0], 1e-5)
    with pytest.raises(NotImplementedError):
        M.grid_sample(img, grid, "zeros", True)
    with pytest.raises(NotImplementedError):   # the sampled image is data on the reference's path
        M.grid_sample(img.clone().requires_grad_(True), grid, "border", True)


@pytest.mark.parametrize("B,H,W", [(2, 24, 32), (3, 192, 640), (1, 3, 3)])
def test_reprojection_loss(M, B, H, W):
    g = torch.Generator(device=DEV).manual_seed(4)
    pred = _rand(g, B, 3, H, W)
    tgt = (pred + 0.1 * torch.randn(B, 3, H, W, generator=g, device=DEV)).clamp(0, 1)
    p1, p2 = pred.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    out = M.ReprojectionLoss()(p1, tgt)
    ref = O.photometric_error(p2, tgt)
    _exact(out, ref)
    cot = torch.randn(ref.shape, generator=g, device=DEV)
    _close(torch.autograd.grad((out * cot).sum(), p1)[0], torch.autograd.grad((ref * cot).sum(), p2)[0], 1e-4)


@pytest.mark.parametrize("B,h,w", [(2, 24, 32), (12, 192, 640), (2, 24, 80)])
def test_smooth_loss(M, B, h, w):
    g = torch.Generator(device=DEV).manual_seed(5)
    disp, color = _rand(g, B, 1, h, w), _rand(g, B, 3, h, w)
    d1, d2 = disp.clone().requires_grad_(True), disp.clone().requires_grad_(True)
    out = M.SmoothLoss()(d1, color)
    ref = O.smoothness(d2, color)
    assert out.shape == ref.shape
    assert abs(float(out.detach()) - float(ref.detach())) <= 2e-6 * abs(float(ref.detach()))
    _close(torch.autograd.grad(out * 3.0, d1)[0], torch.autograd.grad(ref * 3.0, d2)[0], 1e-4)


def test_against_reference_golden(M):
    """The reference's own CPU outputs for each symbol (tests/golden/ops.npz)."""
    z = _golden()
    B, _, H, W = z["bp_depth"].shape
    x = z["up_in"].clone().requires_grad_(True)
    up = M.interpolate(x, H, W, "bilinear", False)
    _close(up, z["up_out"], 1e-6)
    _close(torch.autograd.grad((up * z["up_cot"]).sum(), x)[0], z["up_grad"], 1e-5)
    d = z["d2d_in"].clone().requires_grad_(True)
    s, dep = M.disparity2depth(d, 0.1, 100.0)
    _close(s, z["d2d_scaled"], 1e-6)
    _close(dep, z["d2d_depth"], 1e-6)
    _close(torch.autograd.grad((s * z["d2d_cot0"]).sum() + (dep * z["d2d_cot1"]).sum(), d)[0], z["d2d_grad"], 1e-5)
    dp = z["bp_depth"].clone().requires_grad_(True)
    cam = M.Depth2PointCloud(B, H, W)(dp, z["bp_invK"])
    _close(cam, z["bp_cam"], 1e-6)
    _close(torch.autograd.grad((cam * z["bp_cot"]).sum(), dp)[0], z["bp_grad"], 1e-5)
    c, T = z["pj_cam"].clone().requires_grad_(True), z["pj_T"].clone().requires_grad_(True)
    grid = M.PointCloud2Pixel(B, H, W)(c, z["pj_K"], T)
    _close(grid, z["pj_grid"], 1e-5)
    gc, gT = torch.autograd.grad((grid * z["pj_cot"]).sum(), [c, T])
    _close(gc, z["pj_grad_cam"], 1e-4)
    _close(gT, z["pj_grad_T"], 1e-4)
    gg = z["gs_grid"].clone().requires_grad_(True)
    out = M.grid_sample(z["gs_img"], gg, "border", True)
    _close(out, z["gs_out"], 1e-5)
    _close(torch.autograd.grad((out * z["gs_cot"]).sum(), gg)[0], z["gs_grad"], 1e-4)
    p = z["rp_pred"].clone().requires_grad_(True)
    rep = M.ReprojectionLoss()(p, z["rp_target"])
    _close(rep, z["rp_out"], 1e-5)
    _close(torch.autograd.grad((rep * z["rp_cot"]).sum(), p)[0], z["rp_grad"], 1e-4)
    sd = z["sm_disp"].clone().requires_grad_(True)
    sm = M.SmoothLoss()(sd, z["sm_color"])
    _close(sm, z["sm_out"], 1e-5)
    _close(torch.autograd.grad(sm, sd)[0], z["sm_grad"], 1e-4)


def test_composed_operators_equal_fused_loss(M):
    """processor.py:139-218 written with the symbol-level operators gives the fused kernel's loss and gradients."""
    import md2_b200.synthetic as syn
    from md2_b200 import functional as F_
    B, H, W, frame_ids = 2, 64, 96, [0, -1, 1]
    inputs, outputs = syn.make_batch(B, H, W, frame_ids, 4, 7, "smooth", device=DEV, requires_grad=False,
                                     pose_fn=M.param2matrix)
    noise = syn.make_noise(B, 2, H, W, 4, 7, device=DEV)
    disps = [outputs[("disp", s)].clone().requires_grad_(True) for s in range(4)]
    Ts = [outputs[("c2c", f, 0)].clone().requires_grad_(True) for f in frame_ids[1:]]
    K, inv_K, tgt = inputs[("K", 0)], inputs[("inv_K", 0)], inputs[("color", 0, 0)]
    bp, pj, rl, sl = M.Depth2PointCloud(B, H, W), M.PointCloud2Pixel(B, H, W), M.ReprojectionLoss(), M.SmoothLoss()
    total = 0
    for s in range(4):
        disp = M.interpolate(disps[s], H, W, "bilinear", False)
        _, depth = M.disparity2depth(disp, 0.1, 100.0)
        cam = bp(depth, inv_K)
        rep = [rl(M.grid_sample(inputs[("color", f, 0)], pj(cam, K, Ts[i]), "border", True), tgt)
               for i, f in enumerate(frame_ids[1:])]
        ident = [rl(inputs[("color", f, 0)], tgt) for f in frame_ids[1:]]
        ident = torch.cat(ident, 1) + 0.00001 * noise[s]
        comb = torch.cat([ident, torch.cat(rep, 1)], 1)
        to_opt, _ = torch.min(comb, dim=1)
        total = total + to_opt.mean() + 1e-3 * sl(disps[s], inputs[("color", 0, s)]) / (2 ** s)
    total = total / 4
    grads = torch.autograd.grad(total, disps + Ts)
    d2 = [d.detach().clone().requires_grad_(True) for d in disps]
    T2 = [t.detach().clone().requires_grad_(True) for t in Ts]
    res = F_.view_synthesis_loss(tgt, [inputs[("color", f, 0)] for f in frame_ids[1:]], d2,
                                 [inputs[("color", 0, s)] for s in range(4)], K, inv_K, T2, noise=noise)
    fused = torch.autograd.grad(res["loss"], d2 + T2)
    assert abs(float(total.detach()) - float(res["loss"].detach())) <= 1e-6 * abs(float(res["loss"].detach()))
    for a, b in zip(grads[:4], fused[:4]):
        _close(a, b, 2e-4)
    for a, b in zip(grads[4:], fused[4:]):
        _close(a, b, 2e-3)


def test_mean_inv_depth(M):
    g = torch.Generator(device=DEV).manual_seed(6)
    depth = 0.2 + 20 * _rand(g, 3, 1, 96, 160)
    d1, d2 = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    out = M.mean_inv_depth(d1)
    ref = (1 / d2).mean(3, True).mean(2, True)
    assert out.shape == ref.shape
    _close(out, ref, 1e-6)
    cot = torch.randn(ref.shape, generator=g, device=DEV)
    _close(torch.autograd.grad((out * cot).sum(), d1)[0], torch.autograd.grad((ref * cot).sum(), d2)[0], 1e-5)


def test_compute_posecnn_branch(M):
    """pose_type='posecnn' (processor.py:153-157) through the drop-in object, against the reference's own run
    (tests/golden/posecnn.npz) and the oracle on the GPU."""
    from types import SimpleNamespace
    from md2_b200.compute import compute
    from test_oracle_ops_golden import _load, posecnn_args
    z = _load("posecnn.npz")
    a = posecnn_args(z, DEV)
    B, _, H, W = a["target"].shape
    opt = SimpleNamespace(frame_ids=[0, -1, 1], scales=range(4), height=H, width=W, min_depth=0.1, max_depth=100.0,
                          pose_type="posecnn", use_automasking=True, disp_smoothness=1e-3)
    inputs = {("color", 0, 0): a["target"], ("color", -1, 0): a["sources"][0], ("color", 1, 0): a["sources"][1],
              ("K", 0): a["K"], ("inv_K", 0): a["inv_K"]}
    outputs = {}
    for s in range(4):
        inputs[("color", 0, s)] = a["color_pyr"][s]
        outputs[("disp", s)] = a["disps"][s]
    for f in (-1, 1):
        outputs[("R", f, 0)], outputs[("T", f, 0)] = a["R"][f], a["T"][f]
    c = compute(opt, torch.device(DEV))
    c.image2warping(inputs, outputs, None, noise=a["noise"])
    c.compute_loss(inputs, outputs, None)
    wrt = a["disps"] + [a["R"][-1], a["R"][1], a["T"][-1], a["T"][1]]
    grads = torch.autograd.grad(outputs["loss"], wrt)
    ref_loss = float(z["loss"])
    assert abs(float(outputs["loss"].detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    for s in range(4):
        _close(outputs[("warp_color", 1, s)], z[f"warp{s}"].to(DEV), 1e-4)
        _close(grads[s], z[f"grad_disp{s}"].to(DEV), 1e-3)
    for g_, k in zip(grads[4:], ["grad_R-1", "grad_R1", "grad_T-1", "grad_T1"]):
        _close(g_, z[k].to(DEV), 2e-3)
    # and the oracle evaluated by ATen on this GPU: same loss to fp32 rounding of the reductions
    b = posecnn_args(z, DEV)
    ref = O.view_synthesis_loss(b["target"], b["sources"], b["disps"], b["color_pyr"], b["K"], b["inv_K"], None,
                                noise=b["noise"], posecnn=[(b["R"][f][:, 0], b["T"][f][:, 0], f < 0) for f in (-1, 1)])
    assert abs(float(outputs["loss"].detach()) - float(ref["loss"].detach())) <= 2e-6 * abs(ref_loss)
    rg = torch.autograd.grad(ref["loss"], b["disps"] + [b["R"][-1], b["R"][1], b["T"][-1], b["T"][1]])
    for x, y in zip(grads, rg):
        _close(x, y, 1e-3)


# ----------------------------------------------------------------------------- ReflectionPad2d (decoder Conv3x3)
@pytest.mark.parametrize("shape,pad,cl", [
    ((2, 16, 12, 20), 1, True),               # the decoder's case: channels-last, C % 4 == 0 (float4 path)
    ((2, 16, 12, 20), 1, False),              # contiguous NCHW
    ((3, 6, 7, 9), (2, 1, 3, 0), True),       # channels-last, C % 4 != 0 (scalar path), asymmetric pads
    ((1, 8, 5, 4), (3, 3, 4, 4), True),       # pads of size - 1: every mirror overlaps
    ((2, 3, 2, 2), (1, 1, 1, 1), False),      # 2 x 2 planes: an element is read up to 9 times
    ((1, 1, 6, 6), 2, True),                  # C == 1: channels-last and NCHW coincide
    ((12, 32, 96, 320), 1, True),             # a real decoder level
])
def test_reflection_pad2d_matches_aten(shape, pad, cl):
    """model_layer/depth_decoder.py:36-50 (Conv3x3.pad): forward bit-identical to nn.ReflectionPad2d in both memory
    formats, output in the input's format; backward equal to ATen's (integer-valued gradients make every summation
    order exact, so the comparison is bit-exact)."""
    import md2_b200.modules as M
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g).to(DEV)
    fmt = torch.channels_last if cl else torch.contiguous_format
    x = x.contiguous(memory_format=fmt)
    xa = x.detach().clone().requires_grad_(True)
    xb = x.detach().clone().contiguous(memory_format=fmt).requires_grad_(True)
    ours, ref = M.ReflectionPad2d(pad)(xb), torch.nn.ReflectionPad2d(pad)(xa)
    assert ours.shape == ref.shape and torch.equal(ours, ref)
    assert ours.is_contiguous(memory_format=fmt)
    go = torch.randint(-8, 9, ref.shape, generator=g).float().to(DEV).contiguous(memory_format=fmt)
    ours.backward(go)
    ref.backward(go)
    assert torch.equal(xb.grad, xa.grad)
    assert xb.grad.is_contiguous(memory_format=fmt)
    # real-valued gradients: equal up to the order of at most nine additions
    gr = torch.randn(ref.shape, generator=g).to(DEV)
    (g1,) = torch.autograd.grad(M.ReflectionPad2d(pad)(xb), xb, gr)
    (g2,) = torch.autograd.grad(torch.nn.ReflectionPad2d(pad)(xa), xa, gr)
    assert torch.allclose(g1, g2, rtol=1e-6, atol=1e-6)


def test_reflection_pad2d_rejects_bad_input_and_swaps_into_modules():
    import md2_b200.modules as M
    x = torch.randn(1, 4, 5, 5, device=DEV)
    with pytest.raises(RuntimeError):
        M.ReflectionPad2d(5)(x)                       # pad must be smaller than the axis (torch raises too)
    with pytest.raises(RuntimeError):
        M.ReflectionPad2d(1)(x.cpu())                 # no CPU path
    with pytest.raises(RuntimeError):
        M.ReflectionPad2d(1)(x.double())
    with pytest.raises(RuntimeError):
        M.ReflectionPad2d(1)(x[0])                    # 3-D
    net = torch.nn.Sequential(torch.nn.ReflectionPad2d(1), torch.nn.Conv2d(4, 8, 3),
                              torch.nn.Sequential(torch.nn.ReflectionPad2d((1, 0, 2, 1)), torch.nn.ELU())).to(DEV)
    ref = net(x)
    assert M.use_channels_last_padding(net) == 2
    assert isinstance(net[0], M.ReflectionPad2d) and net[2][0].padding == (1, 0, 2, 1)
    assert torch.equal(net(x), ref)


# ----------------------------------------------------------------------------- MaxPool2d (encoder stem)
@pytest.mark.parametrize("shape,k,s,p,kind", [
    ((2, 64, 24, 40), 3, 2, 1, "relu"),        # the encoders' stem (depth_encoder.py:29) on post-ReLU data: many ties
    ((3, 6, 7, 9), 3, 2, 1, "randn"),          # C % 4 != 0 (scalar path), odd sizes
    ((1, 8, 5, 5), 2, 2, 0, "ties"),           # every window is one repeated value
    ((2, 4, 9, 11), 3, 1, 1, "randn"),         # stride 1: an element is in nine windows
    ((1, 4, 6, 6), 5, 3, 2, "nan"),            # NaN and isolated -inf inputs
    ((12, 64, 96, 320), 3, 2, 1, "relu"),      # the real stem at batch 12
])
def test_maxpool2d_nhwc_matches_aten(shape, k, s, p, kind):
    """Values bit-identical to nn.MaxPool2d; the winner of every window is ATen's (compared through the gradient: with
    integer-valued output gradients every summation order is exact, so the input gradients must be bit-identical)."""
    import md2_b200.modules as M
    g = torch.Generator().manual_seed(9)
    x = torch.randn(*shape, generator=g)
    if kind == "relu":
        x = x.clamp_min(0)
    elif kind == "ties":
        x = torch.round(x)
        x = x[:, :, :1, :1].expand(shape).clone()
    elif kind == "nan":
        x[0, 0, 2, 2] = float("nan")
        x[0, 2, 1, 1] = float("-inf")
        x[0, 2, 4, 3] = float("-inf")
    x = x.to(DEV).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ref = torch.nn.MaxPool2d(k, s, p)(xa)
    ours = M.MaxPool2d(k, s, p)(xb)
    assert ours.shape == ref.shape and ours.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(torch.nan_to_num(ours, nan=123.0), torch.nan_to_num(ref, nan=123.0))
    go = torch.randint(-8, 9, ref.shape, generator=g).float().to(DEV).contiguous(memory_format=torch.channels_last)
    ref.backward(go)
    ours.backward(go)
    assert torch.equal(xb.grad, xa.grad)
    assert xb.grad.is_contiguous(memory_format=torch.channels_last)


def test_maxpool2d_window_without_a_maximum():
    """A window that holds only -inf has no element greater than the running maximum: the gradient goes to the first
    in-bounds element, as in ATen's CPU and NCHW CUDA kernels (its NHWC CUDA kernel sends it to element (0, 0) of the
    plane instead - a quirk this kernel does not copy; unreachable behind a ReLU)."""
    import md2_b200.modules as M
    x = torch.full((1, 4, 6, 6), float("-inf"))
    x[0, 3] = torch.arange(36.0).view(6, 6)
    go = torch.arange(1.0, 1 + 4 * 3 * 3).view(1, 4, 3, 3)
    xa = x.clone().requires_grad_(True)
    torch.nn.MaxPool2d(3, 2, 1)(xa).backward(go)                                   # CPU reference
    xb = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = M.MaxPool2d(3, 2, 1)(xb)
    out.backward(go.to(DEV))
    assert torch.equal(xb.grad.cpu(), xa.grad)


def test_maxpool2d_swaps_into_the_encoder_and_rejects_bad_input():
    import torchvision
    import md2_b200.modules as M
    enc = torchvision.models.resnet18(weights=None).to(DEV).to(memory_format=torch.channels_last).eval()
    x = torch.rand(2, 3, 64, 96, device=DEV).contiguous(memory_format=torch.channels_last)
    stem = lambda m: m.maxpool(m.relu(m.bn1(m.conv1(x))))
    with torch.no_grad():
        ref = stem(enc)
        assert M.use_channels_last_pooling(enc) == 1 and isinstance(enc.maxpool, M.MaxPool2d)
        assert torch.equal(stem(enc), ref)
    with pytest.raises(RuntimeError):
        M.MaxPool2d(3, 2, 1)(x.cpu())
    with pytest.raises(RuntimeError):
        M.MaxPool2d(3, 2, 1)(x.double())
    with pytest.raises(ValueError):
        M.MaxPool2d(3, 2, 2)                                  # padding > kernel / 2, like torch
    with pytest.raises(RuntimeError):
        M.MaxPool2d(5, 1, 0)(torch.rand(1, 4, 3, 3, device=DEV))   # window larger than the image
