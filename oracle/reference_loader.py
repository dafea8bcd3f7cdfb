"""Import the REAL reference team classifiers from /root/reference (build container only).
TEST INFRASTRUCTURE — see oracle/__init__.py.  /root/reference does not exist on the GPU box, so
only tests/golden/make_golden.py and CPU tests guarded by `available()` use this.

Two shims are needed:
  1. a stub ``supervision`` module (team.py:5 / team_hybrid.py:10 import it for type hints only);
  2. ``torchvision.models.mobilenet_v3_small(pretrained=True)`` (team_hybrid.py:24) would download
     weights — it is patched to ``weights=None`` under ``torch.manual_seed(seed)`` so the trunk is a
     seeded random-init module shared with the GPU path.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hockey", "common", "team_hybrid.py"))


def load(seed: int = 0):
    """Returns the module namespace (team_hybrid module, team module) of the real reference."""
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    import torch
    import torchvision

    if "supervision" not in sys.modules:
        sv = types.ModuleType("supervision")
        sv.Detections = type("Detections", (), {})
        sys.modules["supervision"] = sv
    hockey_dir = os.path.join(REFERENCE_ROOT, "hockey")
    if hockey_dir not in sys.path:
        sys.path.insert(0, hockey_dir)

    real = torchvision.models.mobilenet_v3_small

    def seeded_trunk(*args, **kwargs):
        kwargs.pop("pretrained", None)
        kwargs["weights"] = None
        torch.manual_seed(load.seed)
        return real(**kwargs)

    load.seed = seed
    torchvision.models.mobilenet_v3_small = seeded_trunk
    import common.team_hybrid as team_hybrid  # noqa: E402  (reference module)
    import common.team as team                # noqa: E402
    return team_hybrid, team


def make_hybrid(seed: int = 0, device: str = "cpu"):
    """A real reference HybridTeamClassifier with a seeded random-init trunk (stdout silenced)."""
    team_hybrid, _ = load(seed)
    load.seed = seed
    with contextlib.redirect_stdout(io.StringIO()):
        return team_hybrid.HybridTeamClassifier(device=device)
