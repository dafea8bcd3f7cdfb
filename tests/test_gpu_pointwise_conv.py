"""K6: fused pointwise convolution (tcgen05 TF32 GEMM + bias + activation + strided destinations) against a float64
reference.  The tensor core reads TF32 (10 explicit mantissa bits, truncated) operands and accumulates in fp32, so
the rigorous bound per output is 2^-9 * sum_i |x_i w_i| (+ fp32 accumulation); layout bugs are orders above it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def _ref(x2d, w, b, act):
    y = x2d.double() @ w.double().T + b.double()
    if act != "none":
        y = y / (1 + torch.exp(-y))
    return y


def _bound(x2d, w):
    return (x2d.abs().double() @ w.abs().double().T) * 2.0 ** -9 * 1.05 + 1e-6


@pytest.mark.parametrize("n,h,w_,cin,cout", [(1, 25, 40, 96, 96), (2, 33, 31, 192, 96), (1, 40, 75, 576, 192),
                                             (3, 9, 29, 64, 64), (1, 23, 40, 1152, 576), (1, 16, 16, 32, 128), (1, 7, 3, 96, 288)])
@pytest.mark.parametrize("act", ["none", "silu", "silu_fast"])
def test_matches_float64_reference(ctx, n, h, w_, cin, cout, act):
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
    x = torch.randn((n, cin, h, w_), device="cuda", generator=g).contiguous(memory_format=CL)
    w = torch.randn((cout, cin), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    before = ctx.launch_count()
    y = ctx.pointwise_conv(x, w, b, act)
    assert ctx.launch_count() == before + 1
    assert y.shape == (n, cout, h, w_) and y.is_contiguous(memory_format=CL)
    x2d = x.permute(0, 2, 3, 1).reshape(-1, cin)
    got = y.permute(0, 2, 3, 1).reshape(-1, cout).double()
    ref = _ref(x2d, w, b, act)
    bound = _bound(x2d, w) * (1.1 if act == "none" else 1.2)       # |d silu / dv| <= 1.1
    assert bool(((got - ref).abs() <= bound).all()), float(((got - ref).abs() / bound).max())
    assert float((got - ref).abs().max()) > 0 or cin <= 32                # TF32 really is in play
    # and against the library path it replaces (cuDNN TF32 convolution + K5 epilogue): same precision class
    if act != "none":
        lib = ctx.bias_act(torch.conv2d(x, w.view(cout, cin, 1, 1).contiguous(memory_format=CL)), b, act)
        assert float((lib - y).abs().max()) <= 2 * float(bound.max())


def test_strided_and_dual_destinations(ctx):
    """C2f.cv1: all 2c channels into the concat buffer at an offset, the second half also dense."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n, h, w_, cin, c = 2, 46, 80, 96, 48
    x = torch.randn((n, cin, h, w_), device="cuda", generator=g).contiguous(memory_format=CL)
    w = torch.randn((2 * c, cin), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn(2 * c, device="cuda", generator=g)
    plain = ctx.pointwise_conv(x, w, b, "silu_fast")
    cat = torch.full((n, 4 * c + 16, h, w_), -7.0, device="cuda").contiguous(memory_format=CL)
    half = torch.full((n, c, h, w_), -7.0, device="cuda").contiguous(memory_format=CL)
    out = ctx.pointwise_conv(x, w, b, "silu_fast", out1=cat, out1_off=16, out2=half, out2_off=0, c2_begin=c, c2_count=c)
    assert out is cat
    assert torch.equal(cat[:, 16:16 + 2 * c], plain) and torch.equal(half, plain[:, c:])
    assert bool((cat[:, :16] == -7.0).all()) and bool((cat[:, 16 + 2 * c:] == -7.0).all())
    # input that is itself a channel slice of a wider buffer is rejected by the tensor-shape check, not silently misread
    with pytest.raises(Exception):
        ctx.pointwise_conv(cat[:, :96], w, b)


def test_unsupported_channel_counts_fail_loudly(ctx):
    from hvb import HvbError
    x = torch.randn((1, 48, 8, 8), device="cuda").contiguous(memory_format=CL)
    with pytest.raises(HvbError):
        ctx.pointwise_conv(x, torch.randn(96, 48, device="cuda"), torch.zeros(96, device="cuda"))
    x = torch.randn((1, 64, 8, 8), device="cuda").contiguous(memory_format=CL)
    with pytest.raises(HvbError):
        ctx.pointwise_conv(x, torch.randn(80, 64, device="cuda"), torch.zeros(80, device="cuda"))
