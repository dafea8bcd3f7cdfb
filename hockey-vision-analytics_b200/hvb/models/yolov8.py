"""Plain-PyTorch YOLOv8 (n / s / m / l / x) detection network exposing the RAW Detect head tensors.

PyTorch runs the backbone forward only (BASELINE north-star); everything after the three raw head
tensors (DFL decode, confidence gate, NMS, scale_boxes) is done by the K2a kernel.  The layer
graph follows the ultralytics 8.3.148 `yolov8.yaml`: the reference loads such a
model with ``YOLO(path)`` at hockey/main.py:77.  Weights are random-init here (no checkpoints
offline) with the ultralytics Detect bias initialisation.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn

SCALES = {  # depth, width, max_channels
    "n": (0.33, 0.25, 1024),
    "s": (0.33, 0.50, 1024),
    "m": (0.67, 0.75, 768),
    "l": (1.00, 1.00, 512),
    "x": (1.00, 1.25, 512),
}
REG_MAX = 16


def _divisible(x: float, d: int = 8) -> int:
    return int(math.ceil(x / d) * d)


class ConvBnAct(nn.Module):
    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True):
        super().__init__()
        self.cv1 = ConvBnAct(c1, c2, 3, 1)
        self.cv2 = ConvBnAct(c2, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = int(c2 * 0.5)
        self.cv1 = ConvBnAct(c1, 2 * self.c, 1, 1)
        self.cv2 = ConvBnAct((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = ConvBnAct(c1, c_, 1, 1)
        self.cv2 = ConvBnAct(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(self.m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class DetectHead(nn.Module):
    """YOLOv8 Detect (legacy 3x3-conv class branch).  forward() returns the raw per-level tensors
    cat(cv2_i(x_i), cv3_i(x_i)) of shape [B, 64 + nc, H_i, W_i]; box channels are side-major."""

    def __init__(self, nc: int, ch: List[int]):
        super().__init__()
        self.nc = nc
        self.stride = (8, 16, 32)
        c2 = max(16, ch[0] // 4, REG_MAX * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(ConvBnAct(x, c2, 3), ConvBnAct(c2, c2, 3), nn.Conv2d(c2, 4 * REG_MAX, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(ConvBnAct(x, c3, 3), ConvBnAct(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.bias_init()

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)

    def forward(self, feats):
        return [torch.cat((self.cv2[i](x), self.cv3[i](x)), 1) for i, x in enumerate(feats)]


class YOLOv8(nn.Module):
    def __init__(self, scale: str = "n", nc: int = 80):
        super().__init__()
        d, w, mc = SCALES[scale]
        self.scale, self.nc = scale, nc

        def ch(c):
            return _divisible(min(c, mc) * w, 8)

        def n(x):
            return max(round(x * d), 1)

        c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        self.b0 = ConvBnAct(3, c64, 3, 2)
        self.b1 = ConvBnAct(c64, c128, 3, 2)
        self.b2 = C2f(c128, c128, n(3), True)
        self.b3 = ConvBnAct(c128, c256, 3, 2)
        self.b4 = C2f(c256, c256, n(6), True)
        self.b5 = ConvBnAct(c256, c512, 3, 2)
        self.b6 = C2f(c512, c512, n(6), True)
        self.b7 = ConvBnAct(c512, c1024, 3, 2)
        self.b8 = C2f(c1024, c1024, n(3), True)
        self.b9 = SPPF(c1024, c1024, 5)
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.h12 = C2f(c1024 + c512, c512, n(3))
        self.h15 = C2f(c512 + c256, c256, n(3))
        self.h16 = ConvBnAct(c256, c256, 3, 2)
        self.h18 = C2f(c256 + c512, c512, n(3))
        self.h19 = ConvBnAct(c512, c512, 3, 2)
        self.h21 = C2f(c512 + c1024, c1024, n(3))
        self.detect = DetectHead(nc, [c256, c512, c1024])

    def forward(self, x) -> List[torch.Tensor]:
        x = self.b2(self.b1(self.b0(x)))
        p3 = self.b4(self.b3(x))
        p4 = self.b6(self.b5(p3))
        p5 = self.b9(self.b8(self.b7(p4)))
        h12 = self.h12(torch.cat((self.up(p5), p4), 1))
        h15 = self.h15(torch.cat((self.up(h12), p3), 1))
        h18 = self.h18(torch.cat((self.h16(h15), h12), 1))
        h21 = self.h21(torch.cat((self.h19(h18), p5), 1))
        return self.detect([h15, h18, h21])


def build_yolov8(scale: str = "n", nc: int = 80, seed: int = 0) -> YOLOv8:
    """Seeded random-init network in eval mode (BatchNorm running stats at their 0/1 defaults)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = YOLOv8(scale, nc).eval()
    torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def level_shapes(h: int, w: int):
    """Head grid shapes for a letterboxed (h, w) input (both multiples of 32)."""
    return [(h // s, w // s) for s in (8, 16, 32)]


def fuse_conv_bn(model: nn.Module) -> nn.Module:
    """Fold every BatchNorm into its convolution, in place (what ultralytics' ``model.fuse()`` does
    before inference, so the reference runs fused convolutions too).  Returns the model."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    for m in model.modules():
        if isinstance(m, ConvBnAct) and isinstance(m.bn, nn.BatchNorm2d):
            m.conv = fuse_conv_bn_eval(m.conv.eval(), m.bn.eval())
            m.bn = nn.Identity()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
