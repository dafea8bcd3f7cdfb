"""CPU: the restated main-loop drivers (oracle/video_reference.py — hockey/main.py:197-322) on a small synthetic clip
with planted heads, and the host-side Config mirror.  The GPU twin is tests/test_gpu_video.py."""
import numpy as np
import torch

H, W, IMGSZ, NP = 720, 1280, 640, 6


def _key(x_chw) -> bytes:
    return np.ascontiguousarray(np.asarray(x_chw)[:, ::31, ::37]).tobytes()


def _clip(n_frames):
    from hvb.synth import planted_head, rink_clip
    from oracle import ultralytics_restated as ur
    frames, boxes, _, _ = rink_clip(3, n_frames, H, W, NP)
    rng = np.random.default_rng(4)
    table = {}
    for f, b in zip(frames, boxes):
        x = ur.preprocess([ur.letterbox(f, IMGSZ, auto=True)])
        hh, ww = x.shape[2:]
        gain, px, py = ur.scale_boxes_geometry((hh, ww), (H, W))
        gt = b.astype(np.float64) * gain + np.array([px, py, px, py])
        lv = [(hh // s, ww // s) for s in (8, 16, 32)]
        table[_key(x[0])] = [torch.from_numpy(t) for t in
                             planted_head(rng, lv, 2, gt, np.array([0] * (NP - 1) + [1]), dup=1, conf_lo=0.5, conf_hi=0.95)]
    return frames, boxes, table


def test_reference_drivers_on_a_planted_clip():
    from hvb.models import build_trunk
    from oracle.video_reference import VideoReference
    frames, boxes, table = _clip(24)
    ref = VideoReference(lambda x: [t[None] for t in table[_key(x[0].numpy())]], 2, build_trunk(0, calibrate=True),
                         imgsz=IMGSZ, conf=0.4, initialization_stride=4, max_initialization_frames=3)
    out = ref.process_video(list(frames))
    assert len(out) == 24
    assert ref.n_fit_crops == 4 * (NP - 1)                      # frames 0,4,8,12: `i > max_initialization_frames` breaks at the 5th
    last = out[-1]
    assert len(last.detections) == NP and sorted(last.labels).count("Goalie") == 1
    assert list(last.goalie_team_ids) == [2] and last.color_lookup.dtype == np.int32
    assert len(last.color_lookup) == NP and set(last.player_team_ids.tolist()) <= {0, 1}
    # players come first in the merged detections, the goalie last (Detections.merge([players, goalies]))
    assert list(last.detections.class_id) == [0] * (NP - 1) + [1]
    # every kept detection carries a distinct positive tracker id (box sizes are re-drawn per frame in this synthetic
    # clip, so identities may switch; what matters here is that the driver threads the ids through)
    for r in out[-6:]:
        ids = r.detections.tracker_id.tolist()
        assert len(ids) == len(set(ids)) == NP and min(ids) >= 1
    # detections come back at the planted positions (letterbox -> decode -> scale_boxes round trip)
    d = last.detections
    order = np.argsort(d.xyxy[:, 0]); gt = boxes[-1][np.argsort(boxes[-1][:, 0])]
    assert np.abs(d.xyxy[order] - gt).max() < 1.5


def test_config_defaults_are_the_reference_defaults():
    from hvb.video import Config
    c = Config()
    assert (c.detection_imgsz, c.detection_confidence) == (1280, 0.4)
    assert (c.track_activation_threshold, c.lost_track_buffer, c.minimum_matching_threshold, c.frame_rate,
            c.minimum_consecutive_frames) == (0.25, 30, 0.8, 30, 2)
    assert (c.initialization_stride, c.max_initialization_frames, c.min_players_for_selection) == (10, 20, 6)
