"""Seeded MobileNetV3-small trunk (features + avgpool -> 576-d), the deep-feature extractor of
HybridTeamClassifier (reference hockey/common/team_hybrid.py:24-28).

The reference asks torchvision for pretrained weights; there is no network here, so the trunk is
random-init under ``torch.manual_seed(seed)`` and the SAME module object/state is shared between
the oracle and the GPU path.  With default BatchNorm statistics a random-init
trunk emits ~1e-8 features; ``calibrate_bn`` makes it well conditioned with a
seeded train-mode pass so that feature tolerances are meaningful.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def build_trunk(seed: int = 0, calibrate: bool = False) -> nn.Module:
    from torchvision import models
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = models.mobilenet_v3_small(weights=None)
    trunk = nn.Sequential(*list(net.children())[:-1])
    if calibrate:
        calibrate_bn(trunk, seed)
    torch.random.set_rng_state(g)
    trunk.eval()
    for p in trunk.parameters():
        p.requires_grad_(False)
    return trunk


def calibrate_bn(trunk: nn.Module, seed: int = 0, batches: int = 4, batch: int = 64) -> None:
    """Seeded BatchNorm calibration: train-mode forwards on random crops-like tensors with
    cumulative averaging (momentum=None), then back to eval."""
    gen = torch.Generator().manual_seed(seed + 1)
    saved = {}
    for m in trunk.modules():
        if isinstance(m, nn.BatchNorm2d):
            saved[m] = m.momentum
            m.momentum = None
            m.reset_running_stats()
    trunk.train()
    with torch.no_grad():
        for _ in range(batches):
            trunk(torch.randn(batch, 3, 128, 64, generator=gen))
    trunk.eval()
    for m, mom in saved.items():
        m.momentum = mom
