"""K5 parity: the libhvb kernels that replace the element-wise work between the backbone convolutions
(bias + SiLU + residual epilogue, NHWC concat/upsample, the layer-0 stem convolution) against plain
PyTorch fp32 on the same inputs, and the K5-driven YOLOv8 forward against the plain nn.Module forward
(ultralytics graph, hockey/main.py:179-184) with the same weights.

Tolerances (floating point, written here as the north-star asks): epilogue / concat 1e-6 relative
(same fp32 operations, only expf/division rounding may differ from torch's kernels by an ulp); stem
convolution 2e-6 of the output scale (different summation order); whole forward 2e-4 of each head's
scale with TF32 off on both sides."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def nhwc(t):
    return t.cuda().contiguous(memory_format=CL)


@pytest.fixture()
def no_tf32():
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


ACTS = {"none": lambda x: x, "silu": F.silu, "relu": F.relu, "hardswish": F.hardswish, "silu_fast": F.silu}


@pytest.mark.parametrize("c,act", [(48, "silu"), (24, "silu"), (66, "none"), (2, "none"), (7, "relu"), (96, "hardswish"), (48, "silu_fast")])
@pytest.mark.parametrize("with_res", [False, True])
def test_bias_act_in_place(ctx, c, act, with_res):
    g = torch.Generator().manual_seed(c)
    x = torch.randn(3, c, 17, 23, generator=g) * 3
    b = torch.randn(c, generator=g)
    r = torch.randn(3, c, 17, 23, generator=g) if with_res else None
    ref = ACTS[act](x + b.view(1, -1, 1, 1))
    if with_res:
        ref = r + ref
    xd = nhwc(x)
    out = ctx.bias_act(xd, b.cuda(), act, residual=nhwc(r) if with_res else None)
    assert out.data_ptr() == xd.data_ptr()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-6)


def test_bias_act_dual_destination_is_the_c2f_layout(ctx):
    """cv1 of a C2f: all 2c channels into the concat buffer, the second half also dense."""
    g = torch.Generator().manual_seed(0)
    c, nb = 24, 2
    x = torch.randn(2, 2 * c, 20, 36, generator=g)
    b = torch.randn(2 * c, generator=g)
    ref = F.silu(x + b.view(1, -1, 1, 1))
    cat = torch.full((2, (2 + nb) * c, 20, 36), -7.0).cuda().contiguous(memory_format=CL)
    y = torch.empty((2, c, 20, 36)).cuda().contiguous(memory_format=CL)
    ctx.bias_act(nhwc(x), b.cuda(), "silu", out1=cat, out1_off=0, out2=y, out2_off=0, c2_begin=c, c2_count=c)
    torch.testing.assert_close(cat[:, :2 * c].cpu(), ref, rtol=1e-6, atol=1e-6)
    assert (cat[:, 2 * c:] == -7.0).all()                       # the other slices are untouched
    torch.testing.assert_close(y.cpu(), ref[:, c:], rtol=1e-6, atol=1e-6)
    # a Bottleneck output going only into its slice (last Bottleneck of the C2f), with the residual
    r = torch.randn(2, c, 20, 36, generator=g)
    t = torch.randn(2, c, 20, 36, generator=g)
    ctx.bias_act(nhwc(t), b[:c].cuda(), "silu", residual=nhwc(r), out1=cat, out1_off=3 * c)
    torch.testing.assert_close(cat[:, 3 * c:].cpu(), r + F.silu(t + b[:c].view(1, -1, 1, 1)), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(cat[:, :2 * c].cpu(), ref, rtol=1e-6, atol=1e-6)


def test_bias_act_detect_head_layout(ctx):
    """Detect: the 64 box channels and the nc class channels of the two 1x1 convolutions land in one
    [B, 64+nc, H, W] tensor (row pitch 66 floats: exercises the 2-wide and scalar vector paths)."""
    g = torch.Generator().manual_seed(1)
    box, cls = torch.randn(2, 64, 9, 11, generator=g), torch.randn(2, 2, 9, 11, generator=g)
    bb, bc = torch.randn(64, generator=g), torch.randn(2, generator=g)
    head = torch.empty((2, 66, 9, 11)).cuda().contiguous(memory_format=CL)
    ctx.bias_act(nhwc(box), bb.cuda(), "none", out1=head, out1_off=0)
    ctx.bias_act(nhwc(cls), bc.cuda(), "none", out1=head, out1_off=64)
    ref = torch.cat((box + bb.view(1, -1, 1, 1), cls + bc.view(1, -1, 1, 1)), 1)
    assert torch.equal(head.cpu(), ref)                          # a plain fp32 add: bit-exact


def test_bias_act_upsampled_destination_is_upsample_plus_concat(ctx):
    """The neck: Concat([Upsample(2)(p5), p4]) — p5's epilogue writes the 2x2-replicated pixels into the
    concat buffer and its own dense copy into another concat slice, in the same pass."""
    g = torch.Generator().manual_seed(2)
    c5, c4 = 32, 20
    raw = torch.randn(2, c5, 6, 10, generator=g)
    b = torch.randn(c5, generator=g)
    p4 = torch.randn(2, c4, 12, 20, generator=g)
    p5 = F.silu(raw + b.view(1, -1, 1, 1))
    cat12 = torch.zeros((2, c5 + c4, 12, 20)).cuda().contiguous(memory_format=CL)
    cat12[:, c5:] = p4.cuda()
    cat21 = torch.full((2, 8 + c5, 6, 10), 3.0).cuda().contiguous(memory_format=CL)
    ctx.bias_act(nhwc(raw), b.cuda(), "silu", out1=cat21, out1_off=8, out2=cat12, out2_off=0, out2_upsample2=True)
    ref12 = torch.cat([F.interpolate(p5, scale_factor=2, mode="nearest"), p4], 1)
    torch.testing.assert_close(cat12.cpu(), ref12, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(cat21[:, 8:].cpu(), p5, rtol=1e-6, atol=1e-6)
    assert (cat21[:, :8] == 3.0).all()


def test_bias_act_rejects_bad_slices(ctx):
    from hvb._ffi import HvbError
    x = torch.zeros((1, 8, 4, 4)).cuda().contiguous(memory_format=CL)
    small = torch.zeros((1, 4, 4, 4)).cuda().contiguous(memory_format=CL)
    with pytest.raises(HvbError):
        ctx.bias_act(x, None, "silu", out1=small)                # 8 channels do not fit a 4-channel row
    with pytest.raises(ValueError):
        ctx.bias_act(torch.zeros((1, 8, 4, 4)).cuda(), None)     # NCHW-dense input is refused, not reinterpreted


@pytest.mark.parametrize("shifts,chans", [([1, 0], (288, 192)), ([0, 0], (96, 48)), ([0, 0, 0, 0], (24, 24, 24, 24)), ([2, 0, 1], (6, 3, 9))])
def test_concat_upsample(ctx, shifts, chans):
    g = torch.Generator().manual_seed(5)
    h, w = 24, 40
    srcs = [torch.randn(3, c, h >> s, w >> s, generator=g) for c, s in zip(chans, shifts)]
    ref = torch.cat([F.interpolate(t, scale_factor=2 ** s, mode="nearest") if s else t for t, s in zip(srcs, shifts)], 1)
    out = ctx.concat_nhwc([nhwc(t) for t in srcs], shifts)
    assert out.is_contiguous(memory_format=CL)
    assert torch.equal(out.cpu(), ref)                            # pure data movement: bit-exact


@pytest.mark.parametrize("c,hw", [(288, (23, 40)), (128, (20, 20)), (12, (4, 20)), (8, (1, 3)), (64, (50, 60))])
def test_sppf_pool_concat(ctx, c, hw):
    g = torch.Generator().manual_seed(c)
    y0 = torch.randn(3, c, *hw, generator=g)
    m = torch.nn.MaxPool2d(5, 1, 2)
    y1 = m(y0); y2 = m(y1); y3 = m(y2)
    out = ctx.sppf_pool_concat(nhwc(y0))
    assert out.is_contiguous(memory_format=CL)
    assert torch.equal(out.cpu(), torch.cat([y0, y1, y2, y3], 1))       # max is exact: bit-identical


@pytest.mark.parametrize("co,hw", [(16, (64, 96)), (48, (96, 160)), (48, (33, 271)), (32, (32, 32)), (64, (64, 130))])
def test_stem_conv(ctx, no_tf32, co, hw):
    g = torch.Generator().manual_seed(co)
    x = torch.rand(2, 3, *hw, generator=g)
    w = torch.randn(co, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(co, generator=g)
    ref = F.silu(F.conv2d(x.double(), w.double(), b.double(), 2, 1)).float()
    out = ctx.stem_conv(x.cuda(), w.numpy(), b.numpy())
    assert out.shape == ref.shape and out.is_contiguous(memory_format=CL)
    assert (out.cpu() - ref).abs().max() <= 2e-6 * ref.abs().max()


@pytest.mark.parametrize("scale,nc,hw", [("n", 1, (128, 640)), ("m", 2, (96, 160)), ("n", 1, (640, 640))])
def test_fused_forward_matches_module(ctx, no_tf32, scale, nc, hw):
    from hvb.models import build_yolov8
    from hvb.models.fused import FusedYOLOv8
    from hvb.models.yolov8 import fuse_conv_bn
    import copy
    model = build_yolov8(scale, nc, 3)
    ref_model = fuse_conv_bn(copy.deepcopy(model)).cuda()
    run = FusedYOLOv8(model, ctx)
    x = torch.rand(2, 3, *hw, generator=torch.Generator().manual_seed(9)).cuda()
    with torch.no_grad():
        ref = ref_model(x)
    before = ctx.launch_count()
    got = run(x)
    assert ctx.launch_count() - before > 50                       # the glue ran in libhvb
    for r, gt in zip(ref, got):
        assert gt.shape == r.shape
        assert (gt - r).abs().max().item() <= 2e-4 * r.abs().max().item()
    # and without the stem kernel (torch conv + epilogue for layer 0)
    got2 = FusedYOLOv8(model, ctx, stem_kernel=False)(x)
    for r, gt in zip(ref, got2):
        assert (gt - r).abs().max().item() <= 2e-4 * r.abs().max().item()


def test_detector_uses_the_glue_by_default_and_agrees_with_plain_torch(ctx, no_tf32):
    """End to end through Detector (K1a -> forward -> K2a): same detections with and without K5."""
    from hvb import Detector
    from hvb.models import build_yolov8
    from hvb.synth import rink_frame
    rng = np.random.default_rng(3)
    frame = rink_frame(rng, 720, 1280, 8)[0]
    model = build_yolov8("n", 2, 1)
    a = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True)
    b = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, glue=False)
    assert a.runner is not None and b.runner is None
    da, db = a(frame), b(frame)
    assert len(da) > 10 and len(db) > 10
    # borderline candidates may flip at the conf / IoU thresholds; the strongest ones must agree
    k = 10
    ia, ib = np.argsort(-da.confidence)[:k], np.argsort(-db.confidence)[:k]
    np.testing.assert_allclose(da.confidence[ia], db.confidence[ib], rtol=0, atol=1e-5)
    np.testing.assert_allclose(da.xyxy[ia], db.xyxy[ib], rtol=0, atol=1e-2)


def test_detector_cuda_graph_replay_equals_eager(ctx):
    """Detector(cuda_graph=True): K1a + the K5-driven forward + K2a captured once per frame shape and replayed."""
    from hvb import Detector
    from hvb.models import build_yolov8
    from hvb.synth import rink_frame
    rng = np.random.default_rng(8)
    f1, f2 = rink_frame(rng, 720, 1280, 8)[0], rink_frame(rng, 720, 1280, 8)[0]
    model = build_yolov8("n", 2, 1)
    eager = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, cuda_graph=False)
    graph = Detector(model, "cuda:0", imgsz=640, conf=2e-3, fuse=True, channels_last=True, cuda_graph=None)
    assert graph.cuda_graph and not eager.cuda_graph
    for f in (f1, f2, f1):
        a, b = eager(f), graph(f)
        assert len(a) == len(b) > 0
        assert np.array_equal(a.xyxy, b.xyxy) and np.array_equal(a.confidence, b.confidence) and np.array_equal(a.class_id, b.class_id)


def test_fast_silu_error_bound(ctx):
    """act=4 (ex2.approx + rcp.approx) against the float64 SiLU: <= 1e-6 relative wherever the output is not negligible
    (x >= -8, |y| >= 2.7e-3) and <= 1e-8 absolute in the far negative tail, where the fp32 rounding of x*log2(e) is
    amplified by the exponential but y itself is below 3e-3; the exact formula (act=1) stays within 5e-7 relative."""
    x = torch.linspace(-30, 30, 1 << 20).view(1, 4, 512, 512).contiguous(memory_format=CL)
    ref = (x.double() * torch.sigmoid(x.double()))
    for act, bound in (("silu_fast", 1e-6), ("silu", 5e-7)):
        got = ctx.bias_act(x.clone(memory_format=torch.preserve_format).cuda(), None, act).cpu().double()
        err = (got - ref).abs()
        main = (x >= -8) & (x.abs() > 1e-3)
        rel = float((err[main] / ref[main].abs()).max())
        tail = float(err[x < -8].max())
        print(act, "max relative error for x >= -8:", rel, " max absolute error for x < -8:", tail)
        assert rel <= bound and tail <= 1e-8, (act, rel, tail)
