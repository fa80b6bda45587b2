"""The oracle (oracle/oracle_torch.py) against outputs of the reference itself.

The fixtures were produced by tests/golden/make_golden.py, which imports and runs the
unmodified reference in the build container.  On the same torch build and device the
oracle uses the same ATen ops in the same order, so agreement is required to a few ulp;
the tolerances below are written for a different CPU / torch build."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, GOLDEN_DIR, load_golden, max_rel, norm_rel, with_grad
from oracle import oracle_torch as O


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_outputs(name):
    args, ref = load_golden(name)
    out = O.loss_and_grads(**with_grad(args))
    assert max_rel(out["loss"], ref["loss"]) <= 1e-6
    for s in range(4):
        assert max_rel(out["depth"][s], ref["depth"][s]) <= 1e-6
        assert norm_rel(out["grad_disp"][s], ref["grad_disp"][s]) <= 1e-5
    for f, g in enumerate(ref["grad_T"]):
        if g is not None:
            assert norm_rel(out["grad_T"][f], g) <= 1e-5
    if "per_pixel" in ref:
        for s in range(4):
            assert max_rel(out["per_pixel"][s], ref["per_pixel"][s]) <= 1e-6
            assert int((out["argmin"][s].to(torch.uint8) != ref["argmin"][s]).sum()) == 0


def test_oracle_pose_matrix_known_answers():
    z = np.load(f"{GOLDEN_DIR}/pose.npz")
    aa = torch.from_numpy(z["aa"]).requires_grad_(True)
    tr = torch.from_numpy(z["tr"]).requires_grad_(True)
    cot = torch.from_numpy(z["cot"])
    for k, inv in enumerate([False, True]):
        M = O.pose_matrix(aa, tr, invert=inv)
        ga, gt = torch.autograd.grad((M * cot[k]).sum(), [aa, tr])
        assert torch.allclose(M, torch.from_numpy(z[f"M{k}"]), rtol=1e-6, atol=1e-7)
        assert torch.allclose(ga, torch.from_numpy(z[f"grad_aa{k}"]), rtol=1e-5, atol=1e-6)
        assert torch.allclose(gt, torch.from_numpy(z[f"grad_tr{k}"]), rtol=1e-5, atol=1e-6)


def test_oracle_fp64_runs_and_is_close():
    args, ref = load_golden("mono_automask", dtype=torch.float64)
    out = O.view_synthesis_loss(**args)
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
