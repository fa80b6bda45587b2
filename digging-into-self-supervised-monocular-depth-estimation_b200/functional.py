"""Autograd wrappers of the fused kernels (the call a user of the reference makes ends here).

``view_synthesis_loss`` replaces compute.image2warping + compute.compute_loss
(/root/reference/model_tool/processor.py:139-218): one fused pass produces the loss, the
depth maps, the per-pixel minimum / argmin and - when any input requires a gradient - the
gradients wrt every disparity scale and every pose matrix.  ``loss.backward()`` then only
scales the stored gradients by the upstream scalar (the loss is the last node of the
reference's graph, model_train.py:68), so nothing is recomputed or re-read.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ._ext import ext


class _FusedLoss(torch.autograd.Function):
    """inputs: n_disp disparities, then n_T transformation matrices (differentiable);
    everything else is data."""

    @staticmethod
    def forward(ctx, meta, *diff):
        ns, S = meta["ns"], meta["S"]
        disps = [d.detach().contiguous() for d in diff[:ns]]
        Ts = [t.detach().contiguous() for t in diff[ns:ns + S]]
        need_grad = any(ctx.needs_input_grad[1:])
        common = (meta["target"], meta["sources"], disps, meta["color_pyr"], meta["K"], meta["inv_K"], Ts,
                  meta["noise"], meta["seed"], meta["automask"], meta["min_depth"], meta["max_depth"],
                  meta["disp_smoothness"], meta["want_per_pixel"])
        if need_grad:
            r = ext().loss_forward_backward(*common, 1.0, meta["seed_tensor"])
            ctx.grads = r[4:]
        else:
            r = ext().loss_forward(*common, meta["seed_tensor"])
            ctx.grads = None
        loss, per_px, argmin, depth = r[0], r[1], r[2], r[3]
        ctx.mark_non_differentiable(argmin, depth)
        if per_px is not None:
            ctx.mark_non_differentiable(per_px)
        return loss.reshape(()), per_px, argmin, depth

    @staticmethod
    def backward(ctx, g_loss, *_):
        if ctx.grads is None:
            raise RuntimeError("fused loss was run without gradients")
        need = [i for i in range(len(ctx.grads)) if ctx.needs_input_grad[1 + i]]
        # one multi-tensor launch instead of one multiply per gradient
        scaled = dict(zip(need, torch._foreach_mul([ctx.grads[i] for i in need], g_loss)))
        return (None,) + tuple(scaled.get(i) for i in range(len(ctx.grads)))


def view_synthesis_loss(target: torch.Tensor, sources: Sequence[torch.Tensor], disps: Sequence[torch.Tensor],
                        color_pyr: Sequence[torch.Tensor], K: torch.Tensor, inv_K: torch.Tensor,
                        Ts: Sequence[torch.Tensor], *, noise: Optional[Sequence[torch.Tensor]] = None,
                        seed: int = 0, automask: bool = True, min_depth: float = 0.1, max_depth: float = 100.0,
                        disp_smoothness: float = 1e-3, want_per_pixel: bool = False,
                        seed_tensor: Optional[torch.Tensor] = None):
    """Fused multi-scale view-synthesis loss.

    target [B,3,H,W]; sources S x [B,3,H,W]; disps / color_pyr per scale [B,1,h,w] / [B,3,h,w];
    K, inv_K [B,4,4]; Ts S x [B,4,4].  noise: per-scale [B,S,H,W] N(0,1) draws for the auto-mask
    tie-breaker (processor.py:195); None draws them on the device from ``seed`` - or, when ``seed_tensor`` (one int64
    on the device) is given, from the value it holds when the kernel runs: a step captured in a CUDA graph advances
    that tensor itself (``seed_tensor.add_(1)``), since a replay cannot change the ``seed`` argument.
    Returns dict(loss 0-dim, depth [ns,B,1,H,W], argmin [ns,B,H,W] uint8, per_pixel or None)."""
    c = lambda t: t.contiguous()
    meta = dict(ns=len(disps), S=len(sources), target=c(target), sources=[c(s) for s in sources],
                color_pyr=[c(x) for x in color_pyr], K=c(K), inv_K=c(inv_K),
                noise=[c(n) for n in noise] if noise is not None else [], seed=int(seed),
                automask=bool(automask), min_depth=float(min_depth), max_depth=float(max_depth),
                disp_smoothness=float(disp_smoothness), want_per_pixel=bool(want_per_pixel), seed_tensor=seed_tensor)
    if not torch.is_grad_enabled():
        disps = [d.detach() for d in disps]
        Ts = [t.detach() for t in Ts]
    loss, per_px, argmin, depth = _FusedLoss.apply(meta, *disps, *Ts)
    return dict(loss=loss, depth=depth, argmin=argmin, per_pixel=per_px)


def view_synthesis_loss_backward(target, sources, disps, color_pyr, K, inv_K, Ts, argmin, grad_loss, *,
                                 automask=True, min_depth=0.1, max_depth=100.0, disp_smoothness=1e-3):
    """Stand-alone backward from a saved argmin (md2_loss_backward): returns
    (grad_disp list, grad_T list) of ``grad_loss * loss``."""
    c = lambda t: t.detach().contiguous()
    r = ext().loss_backward(c(target), [c(s) for s in sources], [c(d) for d in disps],
                            [c(x) for x in color_pyr], c(K), c(inv_K), [c(t) for t in Ts], bool(automask),
                            float(min_depth), float(max_depth), float(disp_smoothness), argmin, grad_loss)
    return list(r[:len(disps)]), list(r[len(disps):])


class _PoseMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        a, t = axisangle.contiguous(), translation.contiguous()
        ctx.save_for_backward(a, t)
        ctx.invert = bool(invert)
        return ext().pose_forward(a, t, ctx.invert)

    @staticmethod
    def backward(ctx, gM):
        a, t = ctx.saved_tensors
        ga, gt = ext().pose_backward(a, t, ctx.invert, gM.contiguous())
        return ga, gt, None


def param2matrix(axisangle, translation, invert=False):
    """model_layer/warp.py:126-153 as one kernel forward and one backward:
    [N,1,3] axis-angle, [N,1,3] translation -> [N,4,4]."""
    return _PoseMatrix.apply(axisangle, translation, invert)
