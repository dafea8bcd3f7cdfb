"""CPU: which convolutions of the YOLOv8 forward the fused runner hands to K6 (hvb_pointwise_conv) — pure host logic
(`hvb.models.fused._Conv`), no GPU.  The policy is the measured one (profiles/r01_k6_pointwise.md): pointwise layers with
c_out <= 96, plus 192 -> 192; everything else stays on cuDNN + K5."""
import numpy as np
import pytest
import torch

from hvb.models import build_yolov8
from hvb.models.fused import _Conv


def _routed(scale, nc=2):
    m = build_yolov8(scale, nc, 3)
    convs = [_Conv(mod, "cpu") for mod in m.modules() if isinstance(mod, torch.nn.Conv2d)]
    return convs, [(k.cin, k.cout) for k in convs if k.pointwise]


def test_yolov8m_routes_the_six_measured_layers():
    convs, routed = _routed("m")
    assert routed == [(96, 96), (192, 96), (192, 192), (64, 64), (64, 64), (64, 64)]
    # the layers measured slower on K6 inside the forward are NOT routed
    shapes = {(k.cin, k.cout) for k in convs if not k.pointwise}
    assert {(576, 192), (384, 192), (1152, 576), (1152, 384)} <= shapes


@pytest.mark.parametrize("scale", ["n", "s", "m"])
def test_routed_layers_satisfy_the_kernel_constraints(scale):
    convs, routed = _routed(scale, nc=1 if scale == "n" else 2)
    assert routed
    for k in convs:
        if not k.pointwise:
            assert k.w_tf32 is None
            continue
        assert k.cin % 32 == 0 and (k.cout % 96 == 0 or k.cout % 64 == 0) and k.cout <= 192
        assert k.w_tf32.shape == (k.cout, k.cin) and k.w_tf32.is_contiguous() and k.w_tf32.dtype == torch.float32
        bits = k.w_tf32.view(torch.int32)
        assert int((bits & 0x1FFF).abs().max()) == 0                        # exactly representable in TF32
        w = k.w.reshape(k.cout, k.cin)
        err = (k.w_tf32 - w).abs()
        assert bool((err <= w.abs() * 2.0 ** -11 + 1e-45).all())            # rounded to nearest, not truncated


def test_three_by_three_and_strided_convolutions_are_never_routed():
    convs, _ = _routed("m")
    for k in convs:
        if k.w.shape[2:] != (1, 1) or tuple(k.stride) != (1, 1):
            assert not k.pointwise
