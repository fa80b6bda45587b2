"""GPU parity of the depth-metrics kernels (md2_b200.metrics -> include/md2_metrics.h -> csrc/md2_metrics.cu)
against the oracle's restatement of model_metric.py:70-106 evaluated by ATen on the same GPU, and against the
reference's own CPU run (tests/golden/metrics.npz).  Tolerance 1e-5 relative (fp32 means over up to 3 M pixels;
the kernels accumulate in fp64); the masked-pixel count and, through it, the medians are exact."""
import pytest
import torch

from oracle import oracle_torch as O
from test_oracle_ops_golden import metrics_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(B, H, W, density, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    depth = 0.5 + 60 * torch.rand(B, 1, H, W, generator=g, device=DEV)
    gt = torch.zeros(B, 1, 375, 1242, device=DEV)
    hit = torch.rand(B, 1, 375, 1242, generator=g, device=DEV) < density
    gt[hit] = 1.0 + 79 * torch.rand(int(hit.sum()), generator=g, device=DEV)
    return depth, gt


def _check(got, ref, tol=1e-5):
    for name, a, b in zip(("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3"), got, ref):
        a, b = float(a), float(b)
        assert abs(a - b) <= tol * max(abs(b), 1e-3), (name, a, b)


@pytest.mark.parametrize("B,H,W,density", [(1, 48, 160, 0.05), (12, 192, 640, 0.05), (12, 192, 640, 1.0), (3, 320, 1024, 0.3),
                                           (2, 375, 1242, 0.1)])
def test_metrics_match_oracle(B, H, W, density):
    import md2_b200.metrics as M
    depth, gt = _case(B, H, W, density, B + H)
    out = M.depth_metrics(depth, gt)
    ref = O.depth_metrics(depth, gt)
    _check(out[:7], ref)
    assert int(out[7]) == int((gt[:, :, 153:371, 44:1197] > 0).sum())


def test_upsampled_prediction_and_medians_are_exact():
    """Median scaling is the only data-dependent constant: with gt == clamp(upsample(depth)) on the crop the scaled
    prediction equals the ground truth exactly, so every error is 0 and every accuracy 1 - but only if the
    up-sampling, both clamps and both medians are bit-identical to ATen's."""
    import md2_b200.metrics as M
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(3)
    depth = 0.5 + 100 * torch.rand(4, 1, 192, 640, generator=g, device=DEV)
    gt = torch.clamp(F.interpolate(depth, [375, 1242], mode="bilinear", align_corners=False), 1e-3, 80)
    out = M.depth_metrics(depth, gt)
    assert out[:4].abs().max().item() == 0.0
    assert out[4:7].min().item() == 1.0


def test_metrics_match_reference_golden():
    import md2_b200.metrics as M
    depth, gt, ref, n = metrics_golden(DEV)
    out = M.depth_metrics(depth, gt)
    _check(out[:7], ref)
    assert int(out[7]) == n
    tup = M.compute_depth_metric({("depth", 0): gt}, {("depth", 0, 0): depth}, "torch")
    assert len(tup) == 7 and all(t.dim() == 0 and t.is_cuda for t in tup)
    _check(tup, ref)


def test_metrics_edge_cases():
    import md2_b200.metrics as M
    depth = torch.ones(1, 1, 48, 160, device=DEV)
    gt = torch.zeros(1, 1, 375, 1242, device=DEV)
    out = M.depth_metrics(depth, gt)                      # nothing masked: NaN metrics, count 0
    assert int(out[7]) == 0 and bool(torch.isnan(out[:7]).all())
    gt[0, 0, 200, 600] = 7.0                              # a single masked pixel: prediction scaled onto it
    out = M.depth_metrics(depth, gt)
    assert int(out[7]) == 1 and out[:4].abs().max().item() == 0.0 and out[4:7].min().item() == 1.0
    gt[0, 0, 100, 600] = 9.0                              # outside the crop rows: ignored
    assert int(M.depth_metrics(depth, gt)[7]) == 1
    with pytest.raises(RuntimeError):
        M.depth_metrics(depth.cpu(), gt)
    with pytest.raises(RuntimeError):
        M.depth_metrics(depth, gt[:, :, :300])            # crop does not fit
    with pytest.raises(NotImplementedError):
        M.compute_depth_metric({("depth", 0): gt}, {("depth", 0, 0): depth}, "numpy")
