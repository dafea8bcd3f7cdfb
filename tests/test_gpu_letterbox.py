"""K1a / K1b parity: uint8 stage bit-exact vs the real cv2.resize + copyMakeBorder driven by the
restated LetterBox; float stage bit-exact vs numpy/torch preprocess; tile geometry vs the restated
InferenceSlicer offsets; size-independent properties at 4K."""
import numpy as np
import pytest
import torch

from hvb import _ffi
from hvb.synth import random_frames, rink_frame
from oracle import supervision_restated as svr
from oracle import ultralytics_restated as ur

pytestmark = pytest.mark.gpu


def run_plan(ctx, frames, mode, imgsz, slice_wh=(640, 640), overlap=(128, 128), auto=True):
    n, h, w, _ = frames.shape
    plan = ctx.letterbox_plan(n, h, w, mode, imgsz, auto, 32, slice_wh, overlap)
    fd = torch.from_numpy(frames).cuda()
    u8 = plan.run_u8(fd).cpu().numpy()
    f32 = plan.run(fd).cpu().numpy()
    return plan, u8, f32


def tile_u8(plan, u8, t):
    c = plan.classes[t["cls"]]
    per = 3 * int(c["out_h"]) * int(c["out_w"])
    o = int(c["out_offset"]) + int(t["batch_index"]) * per
    return u8[o:o + per].reshape(int(c["out_h"]), int(c["out_w"]), 3)


def tile_f32(plan, f32, t):
    c = plan.classes[t["cls"]]
    per = 3 * int(c["out_h"]) * int(c["out_w"])
    o = int(c["out_offset"]) + int(t["batch_index"]) * per
    return f32[o:o + per].reshape(3, int(c["out_h"]), int(c["out_w"]))


@pytest.mark.parametrize("h,w,imgsz", [(720, 1280, 640), (1080, 1920, 1280), (2160, 3840, 1280), (1080, 1920, 640),
                                        (333, 517, 640), (97, 1000, 320), (480, 640, 1280), (1, 1, 32), (721, 1283, 640)])
def test_whole_frame_letterbox_bit_exact(ctx, h, w, imgsz):
    frames = random_frames(h + w, 2, h, w)
    plan, u8, f32 = run_plan(ctx, frames, _ffi.LB_WHOLE, imgsz)
    for t in plan.tiles:
        ref = ur.letterbox(frames[t["frame"]], imgsz, auto=True)
        got = tile_u8(plan, u8, t)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), (h, w, imgsz, np.abs(got.astype(int) - ref).max())
        ref_f = ur.preprocess([ref])[0]
        assert np.array_equal(tile_f32(plan, f32, t), ref_f)
        g = ur.letterbox_geometry(h, w, imgsz, True)
        gain, px, py = ur.scale_boxes_geometry((g["out_h"], g["out_w"]), (h, w))
        assert t["gain"] == np.float32(gain) and t["pad_x"] == px and t["pad_y"] == py


def test_1080p_geometry_is_736x1280(ctx):
    plan = ctx.letterbox_plan(1, 1080, 1920, _ffi.LB_WHOLE, 1280)
    assert (plan.classes[0]["out_h"], plan.classes[0]["out_w"]) == (736, 1280)
    assert plan.read_bytes == 1080 * 1920 * 3 and plan.write_bytes == 3 * 736 * 1280 * 4


def test_all_256_values_divide_exactly(ctx):
    frame = np.arange(256, dtype=np.uint8).repeat(3 * 32).reshape(1, 32, 256, 3).copy()
    plan, u8, f32 = run_plan(ctx, frame, _ffi.LB_WHOLE, 256)
    ref = ur.preprocess([ur.letterbox(frame[0], 256, auto=True)])[0]
    assert np.array_equal(tile_f32(plan, f32, plan.tiles[0]), ref)


@pytest.mark.parametrize("h,w", [(720, 1280), (1080, 1920), (2160, 3840), (1000, 1500)])
def test_sliced_exact_mode_matches_reference_per_tile(ctx, h, w):
    frames = random_frames(h, 2, h, w)
    plan, u8, f32 = run_plan(ctx, frames, _ffi.LB_SLICE_EXACT, 640)
    offs = svr.generate_offsets((w, h), (640, 640), (0.2, 0.2))
    assert plan.tiles_per_frame == len(offs)
    for t in plan.tiles:
        o = offs[t["tile"]]
        assert (t["src_x"], t["src_y"], t["src_x"] + t["src_w"], t["src_y"] + t["src_h"]) == tuple(o)
        tile = svr.crop_image(frames[t["frame"]], o)
        ref = ur.letterbox(np.ascontiguousarray(tile), 640, auto=True)
        got = tile_u8(plan, u8, t)
        assert got.shape == ref.shape and np.array_equal(got, ref), (t["tile"], ref.shape)
        assert np.array_equal(tile_f32(plan, f32, t), ur.preprocess([ref])[0])
    if (h, w) == (2160, 3840):
        assert plan.tiles_per_frame == 40 and len(plan.classes) == 5      # 28x640^2, 7x128x640, 3x640x256, 640x288, 288x640
        assert sum(int(c["tiles_per_frame"]) * int(c["out_h"]) * int(c["out_w"]) for c in plan.classes) == 12902400
    if (h, w) == (720, 1280):
        assert plan.tiles_per_frame == 6


def test_sliced_uniform_mode(ctx):
    frames = random_frames(9, 1, 1080, 1920)
    plan, u8, f32 = run_plan(ctx, frames, _ffi.LB_SLICE_UNIFORM, 640)
    assert len(plan.classes) == 1 and (plan.classes[0]["out_h"], plan.classes[0]["out_w"]) == (640, 640)
    offs = svr.generate_offsets((1920, 1080), (640, 640), (0.2, 0.2))
    for t in plan.tiles:
        tile = np.ascontiguousarray(svr.crop_image(frames[0], offs[t["tile"]]))
        assert np.array_equal(tile_u8(plan, u8, t), ur.letterbox(tile, 640, auto=False))


def test_4k_chunk_properties(ctx):
    """Full-size properties: interior tiles are exact copies of the frame / 255, and the kernel is
    deterministic across runs."""
    rng = np.random.default_rng(0)
    frames = np.stack([rink_frame(rng, 2160, 3840, 12, 2.0)[0] for _ in range(2)])
    plan, u8, f32 = run_plan(ctx, frames, _ffi.LB_SLICE_EXACT, 640)
    for t in plan.tiles:
        if t["src_w"] == 640 and t["src_h"] == 640:
            src = frames[t["frame"], t["src_y"]:t["src_y"] + 640, t["src_x"]:t["src_x"] + 640]
            assert np.array_equal(tile_u8(plan, u8, t), src)
            ref = (src[..., ::-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255))
            assert np.array_equal(tile_f32(plan, f32, t), ref)
    f32b = plan.run(torch.from_numpy(frames).cuda()).cpu().numpy()
    assert np.array_equal(f32, f32b)
