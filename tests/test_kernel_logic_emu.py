"""Kernel logic on the CPU: the phase functions of csrc/md2_tile.cuh, compiled for the host
(tests/host_emu), against the golden vectors and the oracle.

This is NOT the parity test proper (that is tests/test_gpu_parity.py, on the B200, through
the C ABI); it checks indexing, halos, reflection, tile seams, the analytic backward and
the host logic in the CPU-only suite.  The CPU oracle uses ATen's CPU kernels, whose
rounding differs from the CUDA kernels the code replicates, so per-pixel comparisons use
the fp64 arbiter: err(ours, ref64) <= 2 * err(ref32, ref64)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, GOLDEN_DIR, load_golden, max_rel, norm_rel, with_grad
from host_emu import emu
from oracle import oracle_torch as O
import md2_b200.synthetic as syn


def to64(args):
    cv = lambda v: [t.double() for t in v] if isinstance(v, list) else (v.double() if torch.is_tensor(v) else v)
    return {k: cv(v) for k, v in args.items()}


def synth_args(B, H, W, frame_ids, automask, kind, seed, num_scales=4):
    inputs, outputs = syn.make_batch(B, H, W, frame_ids, num_scales, seed, kind, pose_fn=O.pose_matrix,
                                     requires_grad=False)
    srcs = frame_ids[1:]
    args = dict(target=inputs[("color", 0, 0)], sources=[inputs[("color", f, 0)] for f in srcs],
                disps=[outputs[("disp", s)] for s in range(num_scales)],
                color_pyr=[inputs[("color", 0, s)] for s in range(num_scales)],
                K=inputs[("K", 0)], inv_K=inputs[("inv_K", 0)],
                Ts=[inputs["stereo"] if f == "s" else outputs[("c2c", f, 0)].detach() for f in srcs],
                automask=automask,
                noise=syn.make_noise(B, len(srcs), H, W, num_scales, seed) if automask else None)
    return args


def check_against_oracle(args, out, grad_tol=2e-3, skip_T=()):
    ref32 = O.loss_and_grads(**with_grad(args))
    ref64 = O.loss_and_grads(**with_grad(to64(args)))
    ns = len(args["disps"])
    assert abs(float(out["loss"]) - float(ref64["loss"].detach())) <= 2e-5 * abs(float(ref64["loss"].detach()))
    flips = 0
    for s in range(ns):
        assert max_rel(out["depth"][s], ref32["depth"][s]) <= 1e-6
        mism = out["argmin"][s].long() != ref32["argmin"][s]
        flips += int(mism.sum())
        # every argmin mismatch must be a numerical tie
        if mism.any():
            gap = (out["per_pixel"][s] - ref32["per_pixel"][s]).abs()[mism]
            assert float(gap.max()) <= 1e-4
        e_ours = (out["per_pixel"][s].double() - ref64["per_pixel"][s]).abs().max()
        e_ref = (ref32["per_pixel"][s].double() - ref64["per_pixel"][s]).abs().max()
        assert float(e_ours) <= 2.0 * float(e_ref) + 1e-6, (s, float(e_ours), float(e_ref))
    assert flips <= 1e-3 * ns * out["argmin"][0].numel() + 2
    if "grad_disp" in out and flips == 0:
        for s in range(ns):
            assert norm_rel(out["grad_disp"][s], ref32["grad_disp"][s]) <= grad_tol, s
        for f, g in enumerate(ref32["grad_T"]):
            if f not in skip_T:  # the stereo baseline is data: its (tiny, noisy) gradient is discarded
                assert norm_rel(out["grad_T"][f], g) <= grad_tol, f
    return ref32, ref64, flips


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emu_fused_matches_reference_golden(name):
    args, ref = load_golden(name)
    out = emu.forward_backward(args)
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 5e-6 * abs(float(ref["loss"]))
    flips = 0
    for s in range(4):
        assert max_rel(out["depth"][s], ref["depth"][s]) <= 1e-6
        if "argmin" in ref:
            flips += int((out["argmin"][s] != ref["argmin"][s]).sum())
    assert flips <= 2
    if flips == 0:
        for s in range(4):
            assert norm_rel(out["grad_disp"][s], ref["grad_disp"][s]) <= 1e-3
        for f, g in enumerate(ref["grad_T"]):
            if g is not None:
                assert norm_rel(out["grad_T"][f], g) <= 1e-3


@pytest.mark.parametrize("B,H,W,frame_ids,automask,kind,seed", [
    (2, 48, 96, [0, -1, 1], True, "smooth", 10),      # several tiles, interior seams
    (1, 40, 72, [0, -1, 1], True, "iid", 11),         # partial tiles at the right / bottom edge
    (1, 32, 64, [0, 1], True, "smooth", 12),          # S = 1 with auto-mask
    (1, 32, 96, [0, -1, 1, "s"], False, "iid", 13),   # S = 3, no auto-mask
    (1, 32, 64, [0, -1, 1, "s", 2], True, "smooth", 14),  # S = 4
])
def test_emu_fused_matches_oracle(B, H, W, frame_ids, automask, kind, seed):
    args = synth_args(B, H, W, frame_ids, automask, kind, seed)
    out = emu.forward_backward(args)
    check_against_oracle(args, out, skip_T=[i for i, f in enumerate(frame_ids[1:]) if f == "s"])


def test_emu_two_scales_and_forward_only_equal_fused():
    args = synth_args(1, 48, 64, [0, -1, 1], True, "smooth", 20, num_scales=2)
    fused = emu.forward_backward(args)
    fwd = emu.forward(args)
    assert float(fwd["loss"]) == pytest.approx(float(fused["loss"]), rel=1e-6)
    for k in ("per_pixel", "argmin", "depth"):
        assert torch.equal(fwd[k], fused[k]), k
    check_against_oracle(args, fused)


def test_emu_standalone_backward_equals_fused_and_scales_with_grad_loss():
    args = synth_args(1, 48, 96, [0, -1, 1], True, "smooth", 21)
    fused = emu.forward_backward(args, grad_loss=1.0)
    bwd = emu.backward(args, fused["argmin"], grad_loss=1.0)
    half = emu.forward_backward(args, grad_loss=0.5)
    for s in range(4):
        assert norm_rel(bwd["grad_disp"][s], fused["grad_disp"][s]) <= 1e-6
        assert norm_rel(2 * half["grad_disp"][s], fused["grad_disp"][s]) <= 1e-6
    for f in range(2):
        assert norm_rel(bwd["grad_T"][f], fused["grad_T"][f]) <= 1e-6
        assert norm_rel(2 * half["grad_T"][f], fused["grad_T"][f]) <= 1e-6


def test_emu_device_noise_is_standard_normal_and_masks_static_pixels():
    args = synth_args(1, 32, 64, [0, -1, 1], True, "smooth", 22)
    # identical source and target: identity loss is ~0 everywhere, so the auto-mask must win
    args["sources"] = [args["target"].clone(), args["target"].clone()]
    args["noise"] = None
    out = emu.forward(args)
    assert int((out["argmin"] >= 2).sum()) <= 0.02 * out["argmin"].numel()
    # identity loss is exactly 0 there, so per_pixel is 1e-5 * min(n0, n1): mean of min of two N(0,1) = -1/sqrt(pi)
    m = float(out["per_pixel"][out["argmin"] < 2].mean()) / 1e-5
    assert abs(m + 0.5642) < 0.05


def test_emu_pose_matches_golden():
    z = np.load(f"{GOLDEN_DIR}/pose.npz")
    aa, tr, cot = torch.from_numpy(z["aa"]), torch.from_numpy(z["tr"]), torch.from_numpy(z["cot"])
    for k, inv in enumerate([False, True]):
        M = emu.pose_forward(aa, tr, inv)
        assert torch.allclose(M, torch.from_numpy(z[f"M{k}"]), rtol=1e-5, atol=1e-6)
        ga, gt = emu.pose_backward(aa, tr, inv, cot[k])
        assert torch.allclose(ga.view(-1, 1, 3), torch.from_numpy(z[f"grad_aa{k}"]), rtol=1e-4, atol=1e-5)
        assert torch.allclose(gt.view(-1, 1, 3), torch.from_numpy(z[f"grad_tr{k}"]), rtol=1e-4, atol=1e-5)


def test_emu_tma_style_zero_fill_and_border_patch_equals_reflected_loads(monkeypatch):
    """The device stages the target / source tiles with TMA (zero fill outside the image) and patches the
    reflection halo afterwards; the host emulation reproduces that path when MD2_EMU_TMA is set."""
    args = synth_args(2, 40, 72, [0, -1, 1], True, "iid", 31)
    ref = emu.forward_backward(args)
    monkeypatch.setenv("MD2_EMU_TMA", "1")
    tma = emu.forward_backward(args)
    fwd = emu.forward(args)
    for k in ("per_pixel", "argmin", "depth"):
        assert torch.equal(ref[k], tma[k]), k
        assert torch.equal(ref[k], fwd[k]), k
    for s in range(4):
        assert torch.equal(ref["grad_disp"][s], tma["grad_disp"][s])
