"""K7 — the device ByteTrack (csrc/k7_bytetrack.cu, hvb.tracker.DeviceByteTrack) against the restated supervision tracker
(oracle/bytetrack_restated.py): tracker ids and kept detections identical frame by frame, for the frame-at-a-time call,
for whole chunks stepped in one launch straight from K2a-layout device tensors, and for several clips in one object."""
import numpy as np
import pytest
import torch

from hvb.detections import Detections
from oracle.bytetrack_restated import ByteTrack as RefByteTrack
from test_tracker import INIT, MAIN, synthetic_detections

pytestmark = pytest.mark.gpu


def _dets(xyxy, conf):
    return Detections(xyxy=xyxy.copy(), confidence=conf.copy(), class_id=np.zeros(len(conf), int))


@pytest.mark.parametrize("seed", [0, 3])
@pytest.mark.parametrize("kw", [MAIN, INIT])
def test_frame_at_a_time_equals_restated_bytetrack(ctx, seed, kw):
    from hvb.tracker import DeviceByteTrack
    ref, mine = RefByteTrack(**kw), DeviceByteTrack(**kw)
    before = ctx.launch_count()
    seen = 0
    for f, (xyxy, conf) in enumerate(synthetic_detections(seed, n_frames=70)):
        if f in (20, 21):
            xyxy, conf = xyxy[:0], conf[:0]
        keep, ids = ref.update_with_detections(xyxy.copy(), conf.copy())
        out = mine.update_with_detections(_dets(xyxy, conf))
        assert np.array_equal(out.tracker_id, ids), f
        assert np.array_equal(out.xyxy, xyxy[keep])
        seen = max(seen, ids.max() if len(ids) else 0)
    assert seen >= 8 and ctx.launch_count() - before >= 70       # one bytetrack_kernel launch per frame


def _pack(clips, f0, f1, md):
    """clips[c][f] = (xyxy, conf) -> K2a-layout tensors, clip-major."""
    nc, nf = len(clips), f1 - f0
    xy = np.zeros((nc * nf, md, 4), np.float32); cf = np.zeros((nc * nf, md), np.float32)
    cl = np.zeros((nc * nf, md), np.int32); cnt = np.zeros((nc * nf,), np.int32)
    for c in range(nc):
        for f in range(f0, f1):
            b, s = clips[c][f]
            i = c * nf + (f - f0)
            xy[i, :len(s)], cf[i, :len(s)], cnt[i] = b, s, len(s)
    return [torch.from_numpy(a).cuda() for a in (xy, cf, cl, cnt)]


def test_chunks_of_several_clips_in_one_launch(ctx):
    """8 clips x 16-frame chunks: one launch per chunk; equals 8 independent reference trackers fed frame by frame."""
    from hvb.tracker import DeviceByteTrack
    n_clips, n_frames, chunk, md = 8, 64, 16, 40
    clips = [synthetic_detections(30 + c, n_frames, n_obj=12 + c) for c in range(n_clips)]
    refs = [RefByteTrack(**MAIN) for _ in range(n_clips)]
    trk = DeviceByteTrack(n_clips=n_clips, **MAIN)
    ctx.launch_count(reset=True)
    for f0 in range(0, n_frames, chunk):
        xy, cf, cl, cnt = _pack(clips, f0, f0 + chunk, md)
        row, tid, tc = (t.cpu().numpy() for t in trk.update_chunk_device(xy, cf, cl, cnt))
        for c in range(n_clips):
            for f in range(chunk):
                keep, ids = refs[c].update_with_detections(clips[c][f0 + f][0].copy(), clips[c][f0 + f][1].copy())
                i = c * chunk + f
                assert tc[i] == len(keep), (c, f0 + f)
                assert np.array_equal(row[i, :tc[i]], keep) and np.array_equal(tid[i, :tc[i]], ids), (c, f0 + f)
    assert ctx.launch_count() == n_frames // chunk


def test_mask_and_overflow_transaction(ctx):
    """Class / confidence mask inside the kernel; a negative input count (K2a overflow pending) leaves the state untouched."""
    from hvb.tracker import DeviceByteTrack
    rng = np.random.default_rng(1)
    frames = synthetic_detections(9, n_frames=24, n_obj=14)
    cls = [rng.integers(0, 3, len(c)).astype(np.int32) for _, c in frames]
    ref, trk = RefByteTrack(**MAIN), DeviceByteTrack(**MAIN)
    md = 32
    for f0 in (0, 8, 16):
        xy, cf, cl, cnt = _pack([frames], f0, f0 + 8, md)
        for k in range(8):
            cl[k, :len(cls[f0 + k])] = torch.from_numpy(cls[f0 + k]).cuda()
        seq = trk.next_seq()
        if f0 == 8:                                              # first attempt: the detector flags frame 3 of the chunk
            bad = cnt.clone(); bad[3] = -1
            _, _, tc = trk.update_chunk_device(xy, cf, cl, bad, min_conf=0.4, class_mask=0b11, seq=seq)
            assert (tc.cpu().numpy() == -2).all()
            # a chunk queued behind the rejected one is rejected as well (the clip only accepts chunks in order)
            _, _, tc = trk.update_chunk_device(xy, cf, cl, cnt, min_conf=0.4, class_mask=0b11, seq=seq + 1)
            assert (tc.cpu().numpy() == -2).all()
        row, tid, tc = (t.cpu().numpy() for t in trk.update_chunk_device(xy, cf, cl, cnt, min_conf=0.4, class_mask=0b11, seq=seq))
        for k in range(8):
            b, s = frames[f0 + k]
            m = (s > 0.4) & (cls[f0 + k] < 2)
            idx = np.nonzero(m)[0]
            keep, ids = ref.update_with_detections(b[m].copy(), s[m].copy())
            assert np.array_equal(row[k, :tc[k]], idx[keep]) and np.array_equal(tid[k, :tc[k]], ids), f0 + k
    trk.reset()
    out = trk.update_with_detections(_dets(*frames[0]))
    assert len(out) == 0 or out.tracker_id.min() >= 1            # MAIN: ids only after 2 consecutive frames


def test_capacity_is_reported_not_truncated(ctx):
    from hvb import _ffi
    from hvb.tracker import DeviceByteTrack
    trk = DeviceByteTrack(**INIT)
    n = 300                                                       # 300 disjoint boxes per frame: > 256 live tracks
    gx, gy = np.meshgrid(np.arange(20), np.arange(15))
    xyxy = np.stack([gx.ravel() * 60, gy.ravel() * 60, gx.ravel() * 60 + 40, gy.ravel() * 60 + 40], 1).astype(np.float32)
    conf = np.linspace(0.5, 0.9, n).astype(np.float32)
    with pytest.raises(_ffi.HvbError):
        trk.update_with_detections(_dets(xyxy, conf))
