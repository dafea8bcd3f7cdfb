"""The planted-detection overlay used by bench.py and the GPU pipeline tests (hvb.synth.PlantedOverlay / DeviceOverlay):
its letterbox / slicer geometry equals the oracle's, planted_head is unchanged by the refactoring into planted_entries,
the index_put path (run here on CPU tensors) writes exactly what apply_host writes — for plain and split head layouts,
for chunks with fewer entries than the table capacity and for chunks with none — and the planted candidates decode to
the boxes they were planted for."""
import numpy as np
import torch

from hvb import synth
from oracle import supervision_restated as svr, ultralytics_restated as ur


def test_geometry_equals_the_oracle():
    for (h, w, s) in [(1080, 1920, 1280), (720, 1280, 1280), (720, 1280, 640), (2160, 3840, 1280), (640, 640, 640), (624, 640, 640),
                      (112, 256, 640), (544, 640, 640), (224, 640, 640), (1000, 777, 640), (333, 517, 640)]:
        g = ur.letterbox_geometry(h, w, s, True)
        oh, ow, gain, px, py = synth.letterbox_geometry(h, w, s)
        assert (oh, ow, px, py) == (g["out_h"], g["out_w"], g["left"], g["top"])
        assert (gain, px, py) == ur.scale_boxes_geometry((oh, ow), (h, w))
    for (w, h) in [(1280, 720), (1920, 1080), (3840, 2160), (1000, 700)]:
        assert np.array_equal(synth.slice_offsets(w, h), svr.generate_offsets((w, h), (640, 640), (0.2, 0.2)))


def _boxes(rng, n, h, w):
    cx, cy = rng.uniform(80, w - 80, n), rng.uniform(120, h - 120, n)
    bw, bh = rng.uniform(30, 80, n), rng.uniform(60, 160, n)
    return np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)


def test_planted_candidates_decode_to_their_boxes_and_nms_keeps_one_each():
    rng = np.random.default_rng(0)
    h, w, imgsz, n = 720, 1280, 640, 7
    boxes = [_boxes(rng, n, h, w) for _ in range(3)]
    cls = [np.array([0] * (n - 1) + [1])] * 3
    ov = synth.PlantedOverlay.whole_frame(3, (h, w), imgsz, boxes, cls, nc=2, dup=3)
    oh, ow, *_ = synth.letterbox_geometry(h, w, imgsz)
    for f in range(3):
        assert ov.count(f) == 3 * n
        heads = [torch.full((1, 66, oh // s, ow // s), -9.0) for s in (8, 16, 32)]
        ov.apply_host(heads, f)
        xyxy, conf, c = ur.predict_from_head(heads, 2, (oh, ow), [(h, w)], 0.4)[0]
        assert len(xyxy) == n and sorted(c.tolist()) == [0] * (n - 1) + [1]
        d = np.abs(xyxy[:, None, :] - boxes[f][None]).max(-1).min(1)
        assert d.max() < 6.0                                    # the un-jittered or a jittered copy (<= 0.6 * 4 px / gain)


class Split:
    def __init__(self, box, cls):
        self.box, self.cls = box, cls


def test_index_put_path_equals_apply_host_with_padding_and_empty_chunks():
    rng = np.random.default_rng(1)
    h, w, imgsz, nfr = 720, 1280, 640, 6
    boxes = [_boxes(rng, k, h, w) for k in (5, 2, 0, 4, 0, 0)]       # frames with no boxes: empty tables
    cls = [np.zeros(len(b), int) for b in boxes]
    ov = synth.PlantedOverlay.whole_frame(3, (h, w), imgsz, boxes, cls, nc=2, dup=2)
    oh, ow, *_ = synth.letterbox_geometry(h, w, imgsz)
    lv = [(oh // s, ow // s) for s in (8, 16, 32)]
    schedule = [[0, 1], [2, 3], [4, 5], [1, 0]]                        # chunk 2 has no entry at all
    dov = ov.to_device("cpu", schedule)
    for rep in range(2):
        for chunk in schedule:
            base = [torch.from_numpy(rng.normal(0, 1, (2, 66, a, b)).astype(np.float32)) for a, b in lv]
            want = [t.clone() for t in base]
            for pos, f in enumerate(chunk):
                ov.apply_host(want, f, 0, pos)
            plain = [t.clone() for t in base]
            dov.begin_chunk(2)
            dov(plain)
            for a, b in zip(plain, want):
                assert torch.equal(a, b)
            split = Split([t[:, :64].clone().contiguous(memory_format=torch.channels_last) for t in base],
                          [t[:, 64:].clone().contiguous(memory_format=torch.channels_last) for t in base])
            dov.begin_chunk(2, repeat=True)
            dov(split)
            for l in range(3):
                assert torch.equal(torch.cat([split.box[l], split.cls[l]], 1), want[l])


def test_sliced_overlay_places_boxes_in_every_tile_that_sees_them():
    h, w = 720, 1280
    offs = synth.slice_offsets(w, h)
    box = np.array([[540.0, 100.0, 560.0, 118.0]])                    # inside the overlap of tiles 0 and 1 (x in [512, 640))
    ov = synth.PlantedOverlay.sliced(0, (h, w), 640, [box], [np.zeros(1, int)], nc=1, dup=1)
    tiles = sorted(t for (f, t) in ov.entries)
    assert tiles == [0, 1]
    for t in tiles:
        x0, y0, x1, y1 = offs[t]
        oh, ow, gain, px, py = synth.letterbox_geometry(int(y1 - y0), int(x1 - x0), 640)
        heads = [torch.full((1, 65, oh // s, ow // s), -9.0) for s in (8, 16, 32)]
        ov.apply_host(heads, 0, t)
        xyxy, conf, c = ur.predict_from_head(heads, 1, (oh, ow), [(int(y1 - y0), int(x1 - x0))], 0.4)[0]
        assert len(xyxy) == 1
        assert np.abs(xyxy[0] + np.array([x0, y0, x0, y0]) - box[0]).max() < 1.0
