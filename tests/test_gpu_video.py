"""End to end through the reference's two drivers (hockey/main.py:197-322): hvb.VideoProcessor on the GPU vs
oracle.video_reference.VideoReference on the CPU, on the same synthetic clip and the SAME raw head tensors (a planted
"model" keyed by the letterboxed frame, so both sides decode identical heads — SURVEY.md H6: random-init YOLO emits
nothing above conf 0.4).  Checked per frame: detections kept by ByteTrack (boxes <= 1e-3 px, confidences, classes),
tracker ids, team ids after the temporal vote, goalie team ids, colour lookup and labels; and that the chunked fast
path reproduces frame-at-a-time processing exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, IMGSZ, NP = 720, 1280, 1280, 8


def _key(x_chw) -> bytes:
    return np.ascontiguousarray(np.asarray(x_chw)[:, ::61, ::67]).tobytes()


@pytest.fixture(scope="module")
def clip():
    """36 frames, 8 drifting players (the last one is a goalie), planted heads per frame."""
    from hvb.synth import planted_head, rink_clip
    from oracle import ultralytics_restated as ur
    frames, boxes, _, _ = rink_clip(5, 36, H, W, NP)
    rng = np.random.default_rng(17)
    table = {}
    for f, b in zip(frames, boxes):
        lb = ur.letterbox(f, IMGSZ, auto=True)
        x = ur.preprocess([lb])
        hh, ww = x.shape[2:]
        gain, px, py = ur.scale_boxes_geometry((hh, ww), (H, W))
        gt = b.astype(np.float64) * gain + np.array([px, py, px, py])
        cls = np.array([0] * (NP - 1) + [1])
        lv = [(hh // s, ww // s) for s in (8, 16, 32)]
        table[_key(x[0])] = [torch.from_numpy(t) for t in planted_head(rng, lv, 2, gt, cls, dup=1, conf_lo=0.5, conf_hi=0.95)]
    return frames, table


class PlantedModel(torch.nn.Module):
    """Stands in for the YOLO forward on both sides: raw heads looked up by the (bit-exact) letterboxed input."""
    nc = 2

    def __init__(self, table):
        super().__init__()
        self.table = table

    def forward(self, x):
        outs = [[], [], []]
        for i in range(x.shape[0]):
            for l, t in enumerate(self.table[_key(x[i].cpu().numpy())]):
                outs[l].append(t.to(x.device))
        return [torch.stack(o) for o in outs]


@pytest.fixture(scope="module")
def reference(clip):
    from hvb.models import build_trunk
    from oracle.video_reference import VideoReference
    frames, table = clip
    trunk = build_trunk(0, calibrate=True)
    ref = VideoReference(lambda x: [t[None] for t in table[_key(x[0].numpy())]], 2, trunk, imgsz=IMGSZ, conf=0.4)
    return trunk, ref, ref.process_video(list(frames))


def _same(got, ref):
    d, r = got.detections, ref.detections
    assert len(d) == len(r)
    if len(r):
        assert np.abs(np.asarray(d.xyxy, np.float64) - r.xyxy).max() <= 1e-3
        assert np.abs(d.confidence - r.confidence).max() <= 1e-6
        assert np.array_equal(np.asarray(d.class_id).astype(int), r.class_id)
        assert np.array_equal(np.asarray(d.tracker_id).astype(int), r.tracker_id)
    assert np.array_equal(np.asarray(got.player_team_ids).astype(int), np.asarray(ref.player_team_ids).astype(int))
    assert np.array_equal(got.goalie_team_ids, ref.goalie_team_ids)
    assert np.array_equal(got.color_lookup, ref.color_lookup)
    assert got.labels == ref.labels


@pytest.mark.parametrize("tracker", ["device", "host"])
def test_process_video_matches_the_reference_drivers(ctx, clip, reference, capsys, tracker):
    from hvb import Config, VideoProcessor
    frames, table = clip
    trunk, ref, ref_out = reference
    vp = VideoProcessor(PlantedModel(table), "cuda:0", Config(), trunk=trunk, tracker=tracker)
    out = list(vp.process_video(list(frames)))
    assert "Classifier fitted." in capsys.readouterr().out
    assert len(out) == len(ref_out) == len(frames)
    # initialisation sampled frames 0,10,20,30 (stride 10, at most 21): 7 player crops each
    assert vp.team_classifier.hybrid_classifier.scaler.n_samples_seen_ == ref.n_fit_crops == 4 * (NP - 1)
    for g, r in zip(out, ref_out):
        _same(g, r)
    tracked = [len(r.detections) for r in ref_out]
    assert max(tracked) == NP                                  # once tracks are confirmed every player + the goalie is tracked
    assert any("Goalie" in r.labels for r in ref_out) and any(len(r.player_team_ids) for r in ref_out)


@pytest.mark.parametrize("tracker", ["device", "host"])
def test_chunked_fast_path_equals_frame_at_a_time(ctx, clip, reference, tracker):
    from hvb import Config, VideoProcessor
    frames, table = clip
    trunk, _, ref_out = reference
    vp = VideoProcessor(PlantedModel(table), "cuda:0", Config(), trunk=trunk, tracker=tracker)
    before = ctx.launch_count()
    out = list(vp.process_video_chunked(list(frames), chunk=16))
    assert ctx.launch_count() > before
    assert len(out) == len(frames)
    for g, r in zip(out, ref_out):
        _same(g, r)


def test_overlay_on_a_real_forward_equals_the_reference_drivers(ctx):
    """The benchmark's step (bench.py): a REAL random-init YOLOv8 forward whose head tensors get planted detections
    scattered in on the device (hvb.synth.PlantedOverlay / DeviceOverlay, Detector.head_hook), then K2a -> K7 -> team
    stage through VideoProcessor.process_chunks on device-resident chunks — against the reference drivers on the CPU with
    the same overlay applied to the CPU forward's heads.  The planted anchors carry identical values on both sides and
    the random-init background stays far below conf, so detections, tracker ids and team ids must agree."""
    import copy
    from hvb import Config, VideoProcessor
    from hvb.models import build_trunk, build_yolov8
    from hvb.models.yolov8 import fuse_conv_bn
    from hvb.synth import PlantedOverlay, rink_clip
    from oracle.video_reference import VideoReference
    h, w, imgsz, n_pl, n_frames, chunk = 720, 1280, 640, 9, 24, 8
    frames, boxes, _, _ = rink_clip(11, n_frames, h, w, n_pl)
    cls = [np.array([0] * (n_pl - 1) + [1])] * n_frames
    ov = PlantedOverlay.whole_frame(5, (h, w), imgsz, boxes, cls, nc=2, dup=3)
    model = build_yolov8("n", 2, 0)
    trunk = build_trunk(0, calibrate=True)
    cpu_model = fuse_conv_bn(copy.deepcopy(model)).eval()
    state = {"f": 0}

    def heads_fn(x):
        with torch.no_grad():
            heads = [t.clone() for t in cpu_model(x)]
        ov.apply_host(heads, state["f"])
        return heads

    ref = VideoReference(heads_fn, 2, trunk, imgsz=imgsz, conf=0.4)
    # fit both sides on the same crops: the players of frames 0, 10, 20 as the reference's initialisation samples them
    ref_out = []
    init_ids = [k for k in range(n_frames) if k % 10 == 0]

    class Seq:                                                     # frames with the cursor the CPU heads_fn needs
        def __iter__(self_inner):
            for k, f in enumerate(frames):
                state["f"] = k
                yield f
    ref.initialize_team_classifier(Seq())
    for k, f in enumerate(frames):
        state["f"] = k
        ref_out.append(ref.process_frame(f))

    cfg = Config(detection_imgsz=imgsz)
    vp = VideoProcessor(model, "cuda:0", cfg, trunk=trunk, detector_kwargs=dict(cuda_graph=False))
    vp.detector.head_hook = ov.to_device("cuda:0", [[k] for k in init_ids])
    vp.initialize_team_classifier(list(frames))
    assert vp.team_classifier.hybrid_classifier.scaler.n_samples_seen_ == ref.n_fit_crops == len(init_ids) * (n_pl - 1)
    vp.detector.head_hook = ov.to_device("cuda:0", [list(range(lo, lo + chunk)) for lo in range(0, n_frames, chunk)])
    dev_chunks = [torch.from_numpy(frames[lo:lo + chunk]).cuda() for lo in range(0, n_frames, chunk)]
    out = list(vp.process_chunks(dev_chunks))
    assert len(out) == n_frames
    for g, r in zip(out, ref_out):
        _same(g, r)
    assert max(len(r.detections) for r in ref_out) == n_pl and sum(len(r.player_team_ids) for r in ref_out) > 100
    # the same through the host-frame API with CUDA graphs (the default of the drop-in) and the host tracker
    vp2 = VideoProcessor(model, "cuda:0", cfg, team_classifier=vp.team_classifier, tracker="host")
    vp2.team_classifier.hybrid_classifier.player_history.clear()
    vp2.detector.head_hook = ov.to_device("cuda:0", [list(range(lo, lo + chunk)) for lo in range(0, n_frames, chunk)])
    out2 = list(vp2.process_video_chunked(list(frames), chunk=chunk, initialize=False))
    for g, r in zip(out2, ref_out):
        _same(g, r)
