"""ops.conv2d (the discriminator's convolution with its own autograd structure) against F.conv2d: forward, first-order
gradients, and the R1-penalty pattern -- gradients of ||d out / d x||^2 with respect to weight, bias and x (double backward)."""
import pytest
import torch
import torch.nn.functional as F

import cips3d_b200
from cips3d_b200 import ops


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw,bias", [
    (3, 8, 1, 1, 0, 9, True), (8, 8, 3, 1, 1, 10, True), (8, 16, 3, 2, 0, 11, False), (8, 16, 1, 2, 0, 9, False), (16, 4, 4, 1, 0, 4, True)])
def test_conv2d_matches_f_conv2d_through_second_order(cin, cout, k, stride, pad, hw, bias):
    g = torch.Generator().manual_seed(cin * 100 + k)
    x0 = torch.randn(3, cin, hw, hw, generator=g, dtype=torch.float64)
    w0 = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) * 0.2
    b0 = torch.randn(cout, generator=g, dtype=torch.float64) if bias else None
    res = {}
    for name, fn in (("ref", F.conv2d), ("ours", ops.conv2d)):
        x, w = x0.clone().requires_grad_(), w0.clone().requires_grad_()
        b = b0.clone().requires_grad_() if bias else None
        y = fn(x, w, b, stride, pad)
        out = (y.tanh()).sum()                                   # nonlinear head so that second derivatives are non-trivial
        gx, = torch.autograd.grad(out, x, create_graph=True)     # R1: gradient of the real-image score w.r.t. the image
        pen = gx.square().sum()
        loss = pen + out
        grads = torch.autograd.grad(loss, [x, w] + ([b] if bias else []))
        res[name] = [y.detach(), gx.detach()] + [t.detach() for t in grads]
    for a, r in zip(res["ours"], res["ref"]):
        assert a.shape == r.shape
        assert (a - r).abs().max().item() <= 1e-10 * max(1.0, r.abs().max().item())


def test_discriminator_r1_gradients_match_the_reference_structure():
    """The module stack with ops.conv2d: R1 gradients equal those of the same stack on F.conv2d (CPU, fp64 weights of a tiny D)."""
    from cips3d_b200 import discriminator as D
    torch.manual_seed(0)
    conv = D.EqualConv2d(4, 6, 3, stride=1, padding=1).double()
    x = torch.randn(2, 4, 8, 8, dtype=torch.float64, requires_grad=True)
    y = conv(x)
    gx, = torch.autograd.grad(y.square().sum(), x, create_graph=True)
    gx.square().sum().backward()
    got = (conv.weight.grad.clone(), conv.bias.grad.clone())
    conv.zero_grad()
    x2 = x.detach().clone().requires_grad_()
    y2 = F.conv2d(x2, conv.weight * conv.scale, conv.bias, 1, 1)
    gx2, = torch.autograd.grad(y2.square().sum(), x2, create_graph=True)
    gx2.square().sum().backward()
    assert (got[0] - conv.weight.grad).abs().max().item() < 1e-10
    assert (got[1] - conv.bias.grad).abs().max().item() < 1e-10
