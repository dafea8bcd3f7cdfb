"""ByteTrack drop-in (hvb/tracker.py) vs the restated supervision ByteTrack (oracle/bytetrack_restated.py)
on synthetic clips with drifting boxes, missed detections, low-score detections and clutter.
The host logic is tested on CPU with an injected numpy cost function; the GPU variant uses the K4b
kernel for every cost matrix."""
import numpy as np
import pytest

from hvb.detections import Detections
from hvb.tracker import ByteTrack
from oracle import supervision_restated as svr
from oracle.bytetrack_restated import ByteTrack as RefByteTrack


def numpy_cost(a, b, scores=None):
    c = svr.iou_distance(a, b)
    return svr.fuse_score(c, scores) if scores is not None else c


def synthetic_detections(seed, n_frames=60, n_obj=10):
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(200, 1700, n_obj), rng.uniform(200, 900, n_obj)
    w, h = rng.uniform(40, 110, n_obj), rng.uniform(100, 250, n_obj)
    frames = []
    for f in range(n_frames):
        cx += rng.uniform(-8, 8, n_obj); cy += rng.uniform(-8, 8, n_obj)
        boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1) + rng.normal(0, 1.5, (n_obj, 4))
        conf = rng.uniform(0.42, 0.95, n_obj)
        conf[rng.random(n_obj) < 0.15] = rng.uniform(0.12, 0.24)          # low-score detections (second association)
        keep = rng.random(n_obj) > 0.1                                      # missed detections
        extra = rng.integers(0, 3)                                          # clutter
        eb = np.stack([rng.uniform(0, 1800, extra), rng.uniform(0, 1000, extra)], 1)
        eb = np.hstack([eb, eb + rng.uniform(30, 100, (extra, 2))])
        b = np.vstack([boxes[keep], eb]).astype(np.float32)
        c = np.concatenate([conf[keep], rng.uniform(0.3, 0.6, extra)]).astype(np.float32)
        p = rng.permutation(len(b))
        frames.append((b[p], c[p]))
    return frames


def run_pair(cost_fn, seed, **kw):
    ref = RefByteTrack(**kw)
    mine = ByteTrack(iou_cost=cost_fn, **kw)
    n_ids = 0
    for xyxy, conf in synthetic_detections(seed):
        keep, ids = ref.update_with_detections(xyxy.copy(), conf.copy())
        out = mine.update_with_detections(Detections(xyxy=xyxy.copy(), confidence=conf.copy(), class_id=np.zeros(len(conf), int)))
        assert np.array_equal(out.tracker_id, ids)
        assert np.array_equal(out.xyxy, xyxy[keep])
        n_ids = max(n_ids, ids.max() if len(ids) else 0)
    assert n_ids >= 8
    return n_ids


MAIN = dict(track_activation_threshold=0.25, lost_track_buffer=30, minimum_matching_threshold=0.8, frame_rate=30,
            minimum_consecutive_frames=2)                                   # hockey/main.py:162-168
INIT = dict(track_activation_threshold=0.25, minimum_consecutive_frames=1, frame_rate=30)   # hockey/main.py:207-211


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("kw", [MAIN, INIT])
def test_tracker_host_logic_matches_restated_bytetrack(seed, kw):
    run_pair(numpy_cost, seed, **kw)


def test_empty_frames_and_reset():
    t = ByteTrack(iou_cost=numpy_cost, **MAIN)
    out = t.update_with_detections(Detections.empty())
    assert len(out) == 0 and out.tracker_id.shape == (0,)
    xyxy, conf = synthetic_detections(3, 3)[0]
    for _ in range(3):
        out = t.update_with_detections(Detections(xyxy=xyxy, confidence=conf, class_id=np.zeros(len(conf), int)))
    assert len(out) > 0
    t.reset()
    assert t.frame_id == 0 and not t.tracked


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [MAIN, INIT])
def test_tracker_with_k4b_kernel(ctx, kw):
    launches = ctx.launch_count(reset=True)
    run_pair(None, 5, **kw)
    assert ctx.launch_count() > 100          # every cost matrix came from the K4b kernel


def _numpy_batched(reqs):
    return [numpy_cost(a, b, sc) if len(a) and len(b) else np.zeros((len(a), len(b))) for a, b, sc in reqs]


def _run_multi(batched_cost, n_clips=4, n_frames=40):
    from hvb.tracker import MultiClipByteTrack
    clips = [synthetic_detections(20 + c, n_frames) for c in range(n_clips)]
    multi = MultiClipByteTrack(n_clips, batched_cost=batched_cost, **MAIN)
    singles = [RefByteTrack(**MAIN) for _ in range(n_clips)]
    seen = 0
    for f in range(n_frames):
        dets = [Detections(xyxy=clips[c][f][0].copy(), confidence=clips[c][f][1].copy(), class_id=np.zeros(len(clips[c][f][1]), int))
                for c in range(n_clips)]
        if f == 7:
            dets[1] = Detections.empty()                           # one clip has an empty frame: the others must not care
        outs = multi.update_with_detections(dets)
        for c in range(n_clips):
            xyxy, conf = (clips[c][f] if not (f == 7 and c == 1) else (np.zeros((0, 4), np.float32), np.zeros(0, np.float32)))
            keep, ids = singles[c].update_with_detections(xyxy.copy(), conf.copy())
            assert np.array_equal(outs[c].tracker_id, ids), (f, c)
            assert np.array_equal(outs[c].xyxy, xyxy[keep])
            seen = max(seen, ids.max() if len(ids) else 0)
    assert seen >= 8


def test_multi_clip_lockstep_equals_independent_trackers():
    _run_multi(_numpy_batched)


@pytest.mark.gpu
def test_multi_clip_lockstep_with_batched_k4b(ctx):
    ctx.launch_count(reset=True)
    _run_multi(None, n_clips=8, n_frames=30)
    per_frame = ctx.launch_count() / 30
    assert 1 <= per_frame <= 6, per_frame                # ~five batched K4b launches per frame for all 8 clips (40 when independent)
