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


def test_fused_forward_routes_pointwise_layers_through_k6(ctx):
    """YOLOv8m forward with the <= 192-channel pointwise layers on K6 vs the same runner on cuDNN TF32 + K5, both against
    the fp32 module: K6 must stay in the precision class of the library path it replaces."""
    import copy
    from hvb.models import build_yolov8
    from hvb.models.fused import FusedYOLOv8
    from hvb.models.yolov8 import fuse_conv_bn
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        model = build_yolov8("m", 2, 3)
        x = torch.rand(2, 3, 96, 160, generator=torch.Generator().manual_seed(9)).cuda()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            ref = fuse_conv_bn(copy.deepcopy(model)).cuda()(x)
        off = FusedYOLOv8(model, ctx)                                  # TF32 off at construction: K6 must not be used
        off(x)
        assert not off.use_pointwise and off.pw_launches == 0
        torch.backends.cudnn.allow_tf32 = True
        lib_run, k6_run = FusedYOLOv8(model, ctx, pointwise_kernel=False), FusedYOLOv8(model, ctx)
        assert k6_run.use_pointwise
        lib, k6 = lib_run(x), k6_run(x)
        # C2f.cv1 + C2f.cv2 of b2 (96->96, 192->96), C2f.cv1 of b4 (192->192) and the three 64 -> 64 Detect convolutions;
        # the 576->192 / 384->192 layers stay on cuDNN + K5 (measured faster there)
        assert lib_run.pw_launches == 0 and k6_run.pw_launches == 6
        errs = []
        for r, a, b in zip(ref, lib, k6):
            scale = r.abs().max().item()
            errs.append(((a - r).abs().max().item() / scale, (b - r).abs().max().item() / scale))
        print("relative error vs fp32 per head (cuDNN TF32 path, K6 path):", [(round(a, 6), round(b, 6)) for a, b in errs])
        for e_lib, e_k6 in errs:
            assert e_k6 <= 5 * e_lib + 5e-3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_detector_with_k6_agrees_with_the_library_path(ctx):
    """Through Detector (K1a -> forward -> K2a), CUDA-graph replay included: strongest detections agree."""
    from hvb import Detector
    from hvb.models import build_yolov8
    from hvb.synth import rink_frame
    rng = np.random.default_rng(3)
    frame = rink_frame(rng, 720, 1280, 8)[0]
    model = build_yolov8("m", 2, 1)
    a = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, cuda_graph=False)
    g = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, cuda_graph=True)
    b = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, cuda_graph=False)
    b.runner.use_pointwise = False
    da, dg, db = a(frame), g(frame), b(frame)
    assert a.runner.pw_launches > 0 and b.runner.pw_launches == 0
    assert len(da) == len(dg) and np.array_equal(da.xyxy, dg.xyxy) and np.array_equal(da.confidence, dg.confidence)
    k = min(10, len(da), len(db))
    ia, ib = np.argsort(-da.confidence)[:k], np.argsort(-db.confidence)[:k]
    np.testing.assert_allclose(da.confidence[ia], db.confidence[ib], rtol=0, atol=2e-3)
