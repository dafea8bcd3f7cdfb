"""GPU: Detector.upload — host frames reach the device unchanged through both routes: pageable frames via the threaded
staging copy (hvb_stage_frames) into the persistent pinned buffer, page-locked frames straight from where they lie."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det(ctx):
    from hvb import Detector
    from hvb.models import build_yolov8
    return Detector(build_yolov8("n", 1), "cuda:0", imgsz=640, conf=0.4)


def test_pageable_frames_are_staged(det):
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (90, 160, 3), dtype=np.uint8) for _ in range(5)]
    assert det._upload_pinned(frames, 5, 90, 160) is None                 # pageable: not the direct route
    for _ in range(3):                                                    # both staging buffers get reused
        dev = det.upload(frames)
        assert np.array_equal(dev.cpu().numpy(), np.stack(frames))
    views = [np.ascontiguousarray(f)[:, ::-1] for f in frames]            # strided views: the general copy
    assert np.array_equal(det.upload(views).cpu().numpy(), np.stack(views))
    one = det.upload(frames[0])
    assert one.shape == (1, 90, 160, 3) and np.array_equal(one[0].cpu().numpy(), frames[0])


def test_page_locked_frames_skip_the_staging_copy(det):
    rng = np.random.default_rng(1)
    block = torch.from_numpy(rng.integers(0, 256, (6, 90, 160, 3), dtype=np.uint8)).pin_memory()
    arr = block.numpy()
    consecutive = [arr[k] for k in range(6)]
    direct = det._upload_pinned(consecutive, 6, 90, 160)
    assert direct is not None and np.array_equal(direct.cpu().numpy(), arr)
    shuffled = [arr[k] for k in (3, 0, 5, 1)]                             # pinned but not consecutive: one copy per frame
    direct = det._upload_pinned(shuffled, 4, 90, 160)
    assert direct is not None and np.array_equal(direct.cpu().numpy(), arr[[3, 0, 5, 1]])
    assert np.array_equal(det.upload(consecutive).cpu().numpy(), arr)
