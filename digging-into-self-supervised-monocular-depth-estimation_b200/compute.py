"""L2 drop-in: an object usable as ``trainer.compute`` for the loss part of the step.

Mirrors compute.image2warping / compute.compute_loss of
/root/reference/model_tool/processor.py:139-218 - same method names, same
``(inputs, outputs, setting)`` dict protocol, same ``opt`` fields - but runs the whole
path as one fused CUDA pass.  image2warping launches it (all its inputs exist by then) and
leaves outputs[("depth", 0, s)] and the loss; compute_loss publishes outputs["loss"].
``("warp_color", f, s)`` is never materialised (its only consumer was compute_loss).

``pose_type == "posecnn"`` (processor.py:153-157) makes the pose of every scale a function of that
scale's depth map, which the fused kernel (one pose per source) does not cover.  That branch is
composed from the symbol-level kernels of modules.py instead: the operators are
this package's CUDA kernels, the cat / min / mean glue between them is the reference's own torch code,
and ``("warp_color", f, s)`` is materialised as in the reference.
"""
from __future__ import annotations

import torch

from . import functional as F_


def _base_seed():
    """A per-object base seed for the on-device auto-mask noise: drawn from torch's global CPU generator (the one the
    reference's torch.randn uses, processor.py:195, so torch.manual_seed makes runs repeatable) and offset by the
    distributed rank, so that DDP ranks seeded alike still draw different noise fields."""
    base = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    return base + (rank << 32)


class compute(object):
    def __init__(self, opt, device):
        self.opt = opt
        self.device = device
        self.step = 0
        self.base_seed = _base_seed()
        self._seed_t = None  # device copy of the running seed (lets a CUDA-graph replay advance it)
        # The fused kernel indexes the pyramid by position: level i has size (H >> i, W >> i) and smoothness weight
        # 1 / 2^i, which equals the reference's 2 ** scale (processor.py:212-214) only for scales = 0, 1, 2, ...
        if not self._posecnn() and list(opt.scales) != list(range(len(list(opt.scales)))):
            raise ValueError(f"md2_b200.compute: opt.scales must be [0, 1, ..., n-1] for the fused path, got "
                             f"{list(opt.scales)} (use pose_type='posecnn' to compose the symbol-level operators)")

    def _posecnn(self):
        return getattr(self.opt, "pose_type", "separate") == "posecnn"

    # ---- posecnn: composed from the symbol-level kernels (processor.py:139-163 with the :153-157 branch)
    def _image2warping_composed(self, inputs, outputs):
        from . import modules as M
        opt = self.opt
        B = inputs[("color", 0, 0)].shape[0]
        H, W = opt.height, opt.width
        backproject, project = M.Depth2PointCloud(B, H, W), M.PointCloud2Pixel(B, H, W)
        for s in opt.scales:
            disp = M.interpolate(outputs[("disp", s)], H, W, "bilinear", False)
            _, depth = M.disparity2depth(disp, opt.min_depth, opt.max_depth)
            outputs[("depth", 0, s)] = depth
            cam = backproject(depth, inputs[("inv_K", 0)])
            for f in opt.frame_ids[1:]:
                if f == "s":
                    T = inputs["stereo"]
                    T = T if T.dim() == 3 else T[None].expand(B, 4, 4)
                else:
                    aa, tr = outputs[("R", f, 0)], outputs[("T", f, 0)]
                    T = F_.param2matrix(aa[:, 0], tr[:, 0] * M.mean_inv_depth(depth)[:, 0], invert=(f < 0))
                grid = project(cam, inputs[("K", 0)], T.contiguous())
                outputs[("warp_color", f, s)] = M.grid_sample(inputs[("color", f, 0)], grid, "border", True)
        return inputs, outputs

    def _compute_loss_composed(self, inputs, outputs, noise=None):
        from . import modules as M
        opt = self.opt
        reprojection, smooth = M.ReprojectionLoss(), M.SmoothLoss()
        target = inputs[("color", 0, 0)]
        srcs = list(opt.frame_ids[1:])
        total = 0
        for i, s in enumerate(opt.scales):
            rep = torch.cat([reprojection(outputs[("warp_color", f, s)], target) for f in srcs], 1)
            if opt.use_automasking:
                ident = torch.cat([reprojection(inputs[("color", f, 0)], target) for f in srcs], 1)
                draw = noise[i] if noise is not None else torch.randn(ident.shape, device=ident.device)
                rep = torch.cat((ident + 0.00001 * draw, rep), dim=1)
            to_optimise = rep if rep.shape[1] == 1 else torch.min(rep, dim=1)[0]
            total = total + to_optimise.mean() + \
                opt.disp_smoothness * smooth(outputs[("disp", s)], inputs[("color", 0, s)]) / (2 ** s)
        outputs["loss"] = total / len(opt.scales)
        return outputs

    def image2warping(self, inputs, outputs, setting=None, noise=None):
        if self._posecnn():
            self._noise = noise
            return self._image2warping_composed(inputs, outputs)
        opt = self.opt
        scales = list(opt.scales)
        srcs = list(opt.frame_ids[1:])
        Ts = [inputs["stereo"] if f == "s" else outputs[("c2c", f, 0)] for f in srcs]
        B = inputs[("color", 0, 0)].shape[0]
        Ts = [t if t.dim() == 3 else t[None].expand(B, 4, 4) for t in Ts]
        if self.step == 0 and not torch.cuda.is_current_stream_capturing():
            # once per object: does torch.matmul on this build round like the kernels' table? (warns, never raises)
            from . import selfcheck
            selfcheck.warn_if_not_replicated(B, opt.height, opt.width, inputs[("color", 0, 0)].device)
        seed_t = None
        if noise is None and bool(opt.use_automasking):
            # the running seed lives on the device and is advanced by a (graph-capturable) in-place add, so that a
            # replayed step draws fresh auto-mask noise like the reference's per-step torch.randn (processor.py:195)
            dev = inputs[("color", 0, 0)].device
            if self._seed_t is None or self._seed_t.device != dev:
                self._seed_t = torch.tensor([self.base_seed & (2 ** 62 - 1)], dtype=torch.int64, device=dev)
            else:
                self._seed_t.add_(1)
            seed_t = self._seed_t
        res = F_.view_synthesis_loss(
            inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in srcs],
            [outputs[("disp", s)] for s in scales], [inputs[("color", 0, s)] for s in scales],
            inputs[("K", 0)], inputs[("inv_K", 0)], Ts, noise=noise, seed=self.base_seed + self.step,
            seed_tensor=seed_t,
            automask=bool(opt.use_automasking), min_depth=opt.min_depth, max_depth=opt.max_depth,
            disp_smoothness=opt.disp_smoothness, want_per_pixel=False)
        self.step += 1
        for i, s in enumerate(scales):
            outputs[("depth", 0, s)] = res["depth"][i]
        outputs[("argmin",)] = res["argmin"]
        outputs[("fused_loss",)] = res["loss"]
        return inputs, outputs

    def compute_loss(self, inputs, outputs, setting=None):
        if self._posecnn():
            return self._compute_loss_composed(inputs, outputs, getattr(self, "_noise", None))
        if ("fused_loss",) not in outputs:
            self.image2warping(inputs, outputs, setting)
        outputs["loss"] = outputs.pop(("fused_loss",))
        return outputs
