"""GPU: the convolution backward of the training path on the repo's own kernels (`ops.conv3d_dgrad`, `ops.conv2d_dgrad`,
`cmfb200_conv_wgrad`) against fp64 autograd of the same torch op the reference calls.  Strict fp32 arithmetic: the gate
is rel-L2 < 1e-5 (fp32 FMA chains over up to 1e5 positions), four orders of magnitude below TF32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _rel(a, b):
    return float((a.detach().cpu().double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _autograd(fn, x, w):
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y = fn(x64, w64)
    g = _rand(*y.shape, seed=7)
    y.backward(g.double())
    return g, x64.grad, w64.grad


@pytest.mark.parametrize("B,Cin,Cout,D,H,W,stride", [(1, 32, 32, 4, 10, 40, 1), (2, 64, 32, 3, 9, 33, 1), (1, 64, 64, 4, 8, 20, 1),
                                                    (1, 32, 64, 4, 12, 36, 2), (2, 64, 64, 6, 8, 20, 2),
                                                    # row pitches that are multiples of 16 bytes: the TMA-staged wgrad
                                                    # (the third one walks the four-deep ring twice per CTA)
                                                    (1, 32, 64, 4, 12, 40, 2), (1, 64, 32, 5, 7, 72, 1),
                                                    (2, 32, 32, 12, 24, 128, 1)])
def test_conv3d_backward(B, Cin, Cout, D, H, W, stride):
    from cmf_b200 import ops

    x, w = _rand(B, Cin, D, H, W, seed=1), _rand(Cout, Cin, 3, 3, 3, seed=2) * 0.05
    g, dx, dw = _autograd(lambda a, b: F.conv3d(a, b, None, stride, 1), x, w)
    got_dx = ops.conv3d_dgrad(g.to(DEV), w.to(DEV), stride)
    got_dw = ops.conv_wgrad(x.to(DEV), g.to(DEV), 3, stride)
    assert got_dx.shape == dx.shape and got_dw.shape == dw.shape
    assert _rel(got_dx, dx) < 1e-5 and _rel(got_dw, dw) < 1e-5, (_rel(got_dx, dx), _rel(got_dw, dw))


@pytest.mark.parametrize("B,Cin,Cout,D,H,W", [(1, 64, 64, 2, 5, 18), (2, 64, 32, 3, 6, 10), (1, 64, 32, 3, 6, 20)])
def test_deconv3d_backward(B, Cin, Cout, D, H, W):
    from cmf_b200 import ops

    x, w = _rand(B, Cin, D, H, W, seed=3), _rand(Cin, Cout, 3, 3, 3, seed=4) * 0.05
    g, dx, dw = _autograd(lambda a, b: F.conv_transpose3d(a, b, None, 2, 1, 1), x, w)
    got_dx = ops.conv3d_dgrad(g.to(DEV), w.to(DEV), 2, transposed=True)
    got_dw = ops.conv_wgrad(g.to(DEV), x.to(DEV), 3, 2)  # roles swapped, see include/cmfb200.h
    assert got_dx.shape == dx.shape and got_dw.shape == dw.shape
    assert _rel(got_dx, dx) < 1e-5 and _rel(got_dw, dw) < 1e-5, (_rel(got_dx, dx), _rel(got_dw, dw))


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,stride,dil", [
    (2, 32, 32, 20, 44, 3, 1, 1), (1, 64, 128, 9, 33, 3, 1, 1), (1, 128, 128, 12, 20, 3, 1, 2), (1, 320, 128, 8, 16, 3, 1, 1),
    (2, 32, 32, 16, 40, 3, 2, 1), (1, 32, 64, 12, 36, 3, 2, 1), (1, 32, 64, 12, 36, 1, 2, 1), (1, 64, 128, 7, 19, 1, 1, 1),
    (2, 128, 32, 5, 9, 1, 1, 1), (1, 3, 32, 18, 34, 3, 1, 1), (4, 128, 32, 1, 2, 1, 1, 1),
    # 16-byte row pitches (TMA-staged wgrad): every kernel family, a 3-channel input, a ring that wraps
    (1, 32, 64, 12, 40, 1, 2, 1), (1, 64, 128, 7, 24, 1, 1, 1), (1, 3, 32, 18, 36, 3, 1, 1), (1, 128, 128, 12, 40, 3, 1, 2),
    (4, 32, 32, 128, 256, 3, 1, 1), (2, 32, 64, 64, 136, 3, 2, 1)])
def test_conv2d_backward(B, Cin, Cout, H, W, k, stride, dil):
    from cmf_b200 import ops

    x, w = _rand(B, Cin, H, W, seed=5), _rand(Cout, Cin, k, k, seed=6) * 0.1
    pad = (k // 2) * dil
    g, dx, dw = _autograd(lambda a, b: F.conv2d(a, b, None, stride, pad, dil), x, w)
    got_dw = ops.conv_wgrad(x.to(DEV), g.to(DEV), k, stride, dil)
    assert got_dw.shape == dw.shape and _rel(got_dw, dw) < 1e-5, _rel(got_dw, dw)
    if Cin != 3:  # the stem's input is the image: no input gradient is ever needed
        got_dx = ops.conv2d_dgrad(g.to(DEV), w.to(DEV), x.shape, stride, dil)
        assert got_dx.shape == dx.shape and _rel(got_dx, dx) < 1e-5, _rel(got_dx, dx)
