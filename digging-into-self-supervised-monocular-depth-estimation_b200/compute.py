"""L2 drop-in: an object usable as ``trainer.compute`` for the loss part of the step.

Mirrors compute.image2warping / compute.compute_loss of
/root/reference/model_tool/processor.py:139-218 - same method names, same
``(inputs, outputs, setting)`` dict protocol, same ``opt`` fields - but runs the whole
path as one fused CUDA pass.  image2warping launches it (all its inputs exist by then) and
leaves outputs[("depth", 0, s)] and the loss; compute_loss publishes outputs["loss"].
``("warp_color", f, s)`` is never materialised (its only consumer was compute_loss).
"""
from __future__ import annotations

import torch

from . import functional as F_


class compute(object):
    def __init__(self, opt, device):
        self.opt = opt
        self.device = device
        self.step = 0

    def _check_supported(self):
        if getattr(self.opt, "pose_type", "separate") == "posecnn":
            # processor.py:153-157 rescales the translation by mean(1/depth) per scale, i.e. T depends
            # on the depth map; that variant is outside the fused path
            raise NotImplementedError("pose_type='posecnn' is not supported by the fused loss; "
                                      "use pose_type 'separate' or 'shared'")

    def image2warping(self, inputs, outputs, setting=None, noise=None):
        self._check_supported()
        opt = self.opt
        scales = list(opt.scales)
        srcs = list(opt.frame_ids[1:])
        Ts = [inputs["stereo"] if f == "s" else outputs[("c2c", f, 0)] for f in srcs]
        B = inputs[("color", 0, 0)].shape[0]
        Ts = [t if t.dim() == 3 else t[None].expand(B, 4, 4) for t in Ts]
        res = F_.view_synthesis_loss(
            inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in srcs],
            [outputs[("disp", s)] for s in scales], [inputs[("color", 0, s)] for s in scales],
            inputs[("K", 0)], inputs[("inv_K", 0)], Ts, noise=noise, seed=self.step,
            automask=bool(opt.use_automasking), min_depth=opt.min_depth, max_depth=opt.max_depth,
            disp_smoothness=opt.disp_smoothness, want_per_pixel=False)
        self.step += 1
        for i, s in enumerate(scales):
            outputs[("depth", 0, s)] = res["depth"][i]
        outputs[("argmin",)] = res["argmin"]
        outputs[("fused_loss",)] = res["loss"]
        return inputs, outputs

    def compute_loss(self, inputs, outputs, setting=None):
        if ("fused_loss",) not in outputs:
            self.image2warping(inputs, outputs, setting)
        outputs["loss"] = outputs.pop(("fused_loss",))
        return outputs
