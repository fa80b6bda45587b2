"""Restatement (numpy, loops) of the index logic of the channels-last network operators - test infrastructure only.

csrc/md2_pad.cu and csrc/md2_pool.cu replace ATen's scatter-style backward kernels of nn.ReflectionPad2d
(model_layer/depth_decoder.py:40) and nn.MaxPool2d (model_layer/depth_encoder.py:29) by gathers: every input element
enumerates the output positions that read it.  These functions state that enumeration in plain Python so that the CPU
suite can pin it against torch's own autograd on shapes the GPU tests do not visit; the CUDA kernels use the same
formulas (pad_sources, the window bounds of maxpool_bwd).
"""
import numpy as np


def pad_sources(i, n, lead, trail):
    """Positions of the padded axis that read input index i (md2_pad.cu pad_sources): itself, its mirror in the
    leading pad, its mirror in the trailing pad."""
    out = [i + lead]
    if 1 <= i <= lead:
        out.append(lead - i)
    if n - 1 - trail <= i <= n - 2:
        out.append(lead + 2 * (n - 1) - i)
    return out


def reflection_pad2d_backward(g_out, pad):
    """g_out [N,C,H+pt+pb,W+pl+pr], pad = (left, right, top, bottom) -> g_in [N,C,H,W] by the gather of reflect_pad_bwd."""
    pl, pr, pt, pb = pad
    N, C, Ho, Wo = g_out.shape
    H, W = Ho - pt - pb, Wo - pl - pr
    g_in = np.zeros((N, C, H, W), g_out.dtype)
    for y in range(H):
        ys = pad_sources(y, H, pt, pb)
        for x in range(W):
            xs = pad_sources(x, W, pl, pr)
            acc = np.zeros((N, C), g_out.dtype)
            for a in ys:
                for b in xs:
                    acc = acc + g_out[:, :, a, b]
            g_in[:, :, y, x] = acc
    return g_in


def maxpool2d_forward(x, k, s, p):
    """x [N,C,H,W] -> (out, winner): the scan of maxpool_fwd - rows then columns of the clipped window, a later
    element wins only if greater or NaN; winner = offset inside the unclipped window, dy * k + dx."""
    N, C, H, W = x.shape
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    out = np.empty((N, C, Ho, Wo), x.dtype)
    win = np.empty((N, C, Ho, Wo), np.uint8)
    for yo in range(Ho):
        for xo in range(Wo):
            y0, x0 = yo * s - p, xo * s - p
            ys, xs = max(y0, 0), max(x0, 0)
            ye, xe = min(y0 + k, H), min(x0 + k, W)
            best = np.full((N, C), -np.inf, x.dtype)
            where = np.full((N, C), (ys - y0) * k + (xs - x0), np.uint8)
            for y in range(ys, ye):
                for xx in range(xs, xe):
                    v = x[:, :, y, xx]
                    take = (v > best) | np.isnan(v)
                    best = np.where(take, v, best)
                    where = np.where(take, (y - y0) * k + (xx - x0), where).astype(np.uint8)
            out[:, :, yo, xo] = best
            win[:, :, yo, xo] = where
    return out, win


def maxpool2d_backward(g_out, win, H, W, k, s, p):
    """The gather of maxpool_bwd: every input element adds the output gradients of the windows whose winner it is."""
    N, C, Ho, Wo = g_out.shape
    g_in = np.zeros((N, C, H, W), g_out.dtype)
    for y in range(H):
        ty = y + p - k + 1
        yo_lo, yo_hi = (0 if ty <= 0 else (ty + s - 1) // s), min((y + p) // s, Ho - 1)
        for x in range(W):
            tx = x + p - k + 1
            xo_lo, xo_hi = (0 if tx <= 0 else (tx + s - 1) // s), min((x + p) // s, Wo - 1)
            acc = np.zeros((N, C), g_out.dtype)
            for yo in range(yo_lo, yo_hi + 1):
                for xo in range(xo_lo, xo_hi + 1):
                    me = (y - (yo * s - p)) * k + (x - (xo * s - p))
                    acc = acc + np.where(win[:, :, yo, xo] == me, g_out[:, :, yo, xo], 0)
            g_in[:, :, y, x] = acc
    return g_in
