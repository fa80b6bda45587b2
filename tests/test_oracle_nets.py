"""The gather formulations behind csrc/md2_pad.cu / csrc/md2_pool.cu (oracle/oracle_nets.py) against torch's own
nn.ReflectionPad2d / nn.MaxPool2d autograd on the CPU - including shapes the GPU tests do not visit (pads of size - 1,
windows larger than the stride grid, odd sizes).  Integer-valued gradients make every summation order exact."""
import numpy as np
import pytest
import torch

from oracle import oracle_nets as N


@pytest.mark.parametrize("shape,pad", [((2, 3, 5, 7), (1, 1, 1, 1)), ((1, 2, 4, 4), (3, 3, 3, 3)), ((1, 1, 6, 3), (2, 0, 5, 1)),
                                       ((2, 2, 2, 2), (1, 1, 1, 1)), ((1, 2, 1, 5), (4, 2, 0, 0)), ((1, 1, 7, 9), (0, 0, 0, 0))])
def test_reflection_pad_backward_gather(shape, pad):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g, requires_grad=True)
    out = torch.nn.ReflectionPad2d(pad)(x)
    go = torch.randint(-9, 10, out.shape, generator=g).float()
    out.backward(go)
    assert np.array_equal(N.reflection_pad2d_backward(go.numpy(), pad), x.grad.numpy())


@pytest.mark.parametrize("shape,k,s,p,kind", [((2, 3, 8, 10), 3, 2, 1, "relu"), ((1, 2, 7, 9), 3, 2, 1, "randn"),
                                              ((1, 2, 9, 9), 5, 3, 2, "randn"), ((1, 1, 6, 6), 2, 2, 0, "ties"),
                                              ((1, 2, 5, 7), 3, 1, 1, "relu"), ((1, 1, 6, 6), 4, 2, 2, "nan"),
                                              ((1, 1, 4, 4), 3, 2, 1, "neginf")])
def test_maxpool_winners_and_backward_gather(shape, k, s, p, kind):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(*shape, generator=g)
    if kind == "relu":
        x = x.clamp_min(0)
    elif kind == "ties":
        x = torch.round(x)[:, :, :1, :1].expand(shape).clone()
    elif kind == "nan":
        x[0, 0, 2, 3] = float("nan")
        x[0, 0, 0, 0] = float("-inf")
    elif kind == "neginf":
        x[:] = float("-inf")       # no element is greater than the running maximum: the first in-bounds one is kept
    xt = x.clone().requires_grad_(True)
    ref = torch.nn.MaxPool2d(k, s, p)(xt)
    out, win = N.maxpool2d_forward(x.numpy(), k, s, p)
    assert np.array_equal(np.nan_to_num(out, nan=7.0, neginf=-7.0), np.nan_to_num(ref.detach().numpy(), nan=7.0, neginf=-7.0))
    go = torch.randint(-9, 10, ref.shape, generator=g).float()
    ref.backward(go)
    got = N.maxpool2d_backward(go.numpy(), win, shape[2], shape[3], k, s, p)
    assert np.array_equal(got, xt.grad.numpy())
