"""End-to-end parity of the drop-in classes against the reference's own outputs (golden fixtures made
by importing hockey/common/team_hybrid.py + team.py) and against the CPU restatement: features,
scaler, affinity, predict labels with the temporal vote, the failure cascade, and the slicer."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import golden_crops  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "team_reference.npz"))


@pytest.fixture(scope="module")
def data():
    frames, crops, labels, positions, tids = golden_crops()
    return dict(frames=frames, crops=crops, labels=labels, positions=positions, tids=tids, n_fit=int((labels >= 0).sum()))


def rel_rowmax(a, b):
    return np.abs(a - b) / (np.abs(b).max(axis=1, keepdims=True) + 1e-300)


def test_hybrid_features_fit_predict_match_reference(ctx, data):
    from hvb import HybridTeamClassifier
    from hvb.models import build_trunk
    clf = HybridTeamClassifier(device="cuda:0", trunk=build_trunk(0), affinity_mode=1)
    n_fit = data["n_fit"]
    feats = clf.extract_all_features(data["crops"])
    assert feats.shape == (47, 625) and feats.dtype == np.float64
    assert rel_rowmax(feats[:, :576], GOLD["deep"].astype(np.float64)).max() <= 1e-3
    np.testing.assert_allclose(feats[:, 576:], GOLD["color"], rtol=1e-9, atol=1e-12)

    clf.fit(data["crops"][:n_fit])
    np.testing.assert_allclose(clf.scaler.mean_[576:], GOLD["scaler_mean"][576:], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(clf.scaler.scale_[576:], GOLD["scaler_scale"][576:], rtol=1e-9, atol=1e-12)
    assert clf.affinity_matrix_.shape == (n_fit, n_fit) and (np.diag(clf.affinity_matrix_) == 1).all()
    assert set(np.unique(clf.cluster_labels)) <= {0, 1}

    per = n_fit // 4
    preds = np.concatenate([clf.predict(data["crops"][f * per:(f + 1) * per], data["tids"][f * per:(f + 1) * per]) for f in range(4)])
    assert preds.dtype.kind == "i" and np.array_equal(preds, GOLD["predict"])
    clf.player_history.clear()
    assert np.array_equal(clf.predict(data["crops"][:n_fit]), GOLD["predict_no_ids"])
    assert clf.predict([]).shape == (0,)
    with pytest.raises(ValueError):
        HybridTeamClassifier(device="cuda:0", trunk=clf.feature_extractor).fit(data["crops"][:3])


def test_predict_from_frame_equals_crop_list_path(ctx, data):
    from hvb import HybridTeamClassifier
    from hvb.models import build_trunk
    from hvb.synth import rink_clip
    frames, boxes, teams, _ = rink_clip(7, 4, 1080, 1920, 10)
    clf = HybridTeamClassifier(device="cuda:0", trunk=build_trunk(0), affinity_mode=1)
    clf.fit(data["crops"][:data["n_fit"]])
    fd = torch.from_numpy(frames).cuda()
    from oracle.supervision_restated import crop_image
    for f in range(2):
        crops = [crop_image(frames[f], b) for b in boxes[f]]
        a = clf.predict(crops)
        b = clf.predict_from_frame(fd, torch.from_numpy(boxes[f]), torch.full((len(boxes[f]),), f, dtype=torch.int32).cuda())
        assert np.array_equal(a, b)


def test_team_classifier_router_and_cascade(ctx, data, capsys):
    from hvb import TeamClassifier
    from hvb.models import build_trunk
    tc = TeamClassifier(device="cuda:0", trunk=build_trunk(0))
    assert tc.use_hybrid and not tc.use_segmentation and not tc.use_robust
    tc.fit(data["crops"][:data["n_fit"]], positions=data["positions"][:data["n_fit"]])
    assert tc.hybrid_classifier.features_normalized_.shape[1] == 625          # positions dropped (team.py:193)
    assert np.array_equal(tc.predict(data["crops"][:10]), GOLD["predict_no_ids"][:10])
    assert tc.predict([]).shape == (0,)
    tc.set_team_names({0: "Away"})
    assert tc.get_team_name(0) == "Away" and tc.get_team_name(5) == "Team 5" and tc.get_segmentation_masks([1]) is None
    # failure cascade: too few crops -> ValueError inside hybrid -> permanent downgrade to the simple rule
    tc2 = TeamClassifier(device="cuda:0", trunk=build_trunk(0))
    tc2.fit(data["crops"][:3])
    assert not tc2.use_hybrid and "Falling back to simple classifier" in capsys.readouterr().out
    assert np.array_equal(tc2.predict(data["crops"][:data["n_fit"]], data["tids"][:data["n_fit"]]), GOLD["simple_predict"])


def test_slicer_device_path_matches_restated_inference_slicer(ctx):
    """4K sliced detection: device path (K1b -> heads -> K2a -> gather -> K2b) vs the restated
    InferenceSlicer driven by the restated ultralytics post-process on the SAME head tensors."""
    from hvb import B200InferenceSlicer, Detector, _ffi
    from hvb.synth import planted_head
    from oracle import supervision_restated as svr
    from oracle import ultralytics_restated as ur

    class PlantedModel(torch.nn.Module):
        """Stands in for the YOLO forward: returns planted heads keyed by the input tile's shape/mean."""
        nc = 1

        def __init__(self):
            super().__init__()
            self.rng = np.random.default_rng(0)
            self.cache = {}

        def heads_for(self, key, hw):
            if key not in self.cache:
                H, W = hw
                lv = [(H // s, W // s) for s in (8, 16, 32)]
                k = 5
                cx, cy = self.rng.uniform(10, W - 10, k), self.rng.uniform(10, H - 10, k)
                sz = self.rng.uniform(8, 30, k)
                gt = np.stack([cx - sz, cy - sz, cx + sz, cy + sz], 1)
                # scores must be distinct ACROSS tiles too: supervision's np.flip(argsort) has no defined tie order
                lo, hi = 0.45 + self.rng.uniform(0, 0.02), 0.95 + self.rng.uniform(0, 0.02)
                self.cache[key] = [torch.from_numpy(t) for t in
                                   planted_head(self.rng, lv, 1, gt, np.zeros(k, int), dup=2, conf_lo=lo, conf_hi=hi)]
            return self.cache[key]

        def forward(self, x):
            outs = [[], [], []]
            for i in range(x.shape[0]):
                key = (tuple(x.shape[2:]), x[i, :, ::61, ::67].cpu().numpy().tobytes())
                for l, t in enumerate(self.heads_for(key, x.shape[2:])):
                    outs[l].append(t.to(x.device))
            return [torch.stack(o) for o in outs]

    from hvb.synth import rink_frame
    rng = np.random.default_rng(3)
    frames = np.stack([rink_frame(rng, 2160, 3840, 12, 2.0)[0] for _ in range(2)])
    model = PlantedModel()
    det = Detector(model, "cuda:0", imgsz=640, conf=0.4)
    slicer = B200InferenceSlicer(detector=det, slice_wh=(640, 640), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.1)
    got = slicer.run_batch(frames)
    # the sync-free device path (full-capacity outputs, no host round trip) agrees with the exact-size one
    fd = torch.from_numpy(frames).cuda()
    a, b = slicer.run_device(fd, sync=True), slicer.run_device(fd, sync=False)
    tot = int(a[4][-1])
    assert int(b[4][-1]) == tot and torch.equal(a[4], b[4]) and int(b[5].min()) >= 0
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x[:tot], y[:tot])

    def callback(tile):
        lb = ur.letterbox(np.ascontiguousarray(tile), 640, auto=True)
        x = torch.from_numpy(ur.preprocess([lb]))
        key = (tuple(x.shape[2:]), x[0, :, ::61, ::67].numpy().tobytes())
        heads = [t[None] for t in model.heads_for(key, x.shape[2:])]
        return ur.predict_from_head(heads, 1, tuple(x.shape[2:]), [tile.shape[:2]], 0.4)[0]

    for f in range(2):
        rx, rc, rk = svr.run_slicer(frames[f], callback, (640, 640), (0.2, 0.2), None, 0.1)
        assert len(got[f]) == len(rx) > 0
        assert np.abs(got[f].xyxy - rx).max() <= 1e-3
        assert np.abs(got[f].confidence - rc).max() <= 1e-6
        assert np.array_equal(got[f].class_id, rk)
    # compat path: user callback on host views, merge through K2b
    from hvb.detections import Detections
    def sv_callback(tile):
        xyxy, conf, cls = callback(tile)
        return Detections(xyxy=xyxy, confidence=conf, class_id=cls)

    compat = B200InferenceSlicer(callback=sv_callback, slice_wh=(640, 640), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.1)
    c = compat(frames[0])
    assert np.abs(c.xyxy - got[0].xyxy).max() <= 1e-3


def test_detector_whole_frame_matches_restated_path(ctx):
    from hvb import Detector
    from hvb.models import build_yolov8
    from oracle import ultralytics_restated as ur
    from hvb.synth import rink_frame
    model = build_yolov8("n", 2, seed=0)
    det = Detector(model, "cuda:0", imgsz=1280, conf=1e-3)          # low conf: random-init emits nothing at 0.4
    rng = np.random.default_rng(1)
    frame = rink_frame(rng, 1080, 1920, 12)[0]
    got = det(frame)
    x = torch.from_numpy(ur.preprocess([ur.letterbox(frame, 1280, auto=True)])).cuda()
    heads = [h.cpu() for h in det.forward_heads(x)]
    rx, rc, rk = ur.predict_from_head(heads, 2, (736, 1280), [(1080, 1920)], 1e-3)[0]
    assert len(got) == len(rx)
    if len(rx):
        assert np.abs(got.confidence - rc).max() <= 1e-6 and np.abs(got.xyxy - rx).max() <= 1e-3
    assert len(det.detect_players(frame)) <= len(got)


def test_pipelined_stream_equals_chunk_api(ctx):
    """HotPath.process_stream (H2D of chunk i+1 overlapped with chunk i) returns what process_chunk returns."""
    from hvb.pipeline import HotPath
    from hvb.synth import rink_frame
    path = HotPath("cuda:0", "n", 2, 640, 0.4, seed=0)
    rng = np.random.default_rng(0)
    chunks = []
    for _ in range(3):
        fr, bx, fi = [], [], []
        for i in range(2):
            f, b, _, _ = rink_frame(rng, 540, 960, 6, 0.5)
            fr.append(f); bx.append(b); fi.append(np.full(len(b), i, np.int32))
        chunks.append((np.stack(fr), np.concatenate(bx).astype(np.float32), np.concatenate(fi)))
    fd = torch.from_numpy(chunks[0][0]).cuda()
    path.fit_from_frames(fd, torch.from_numpy(chunks[0][1]).cuda(), torch.from_numpy(chunks[0][2]).cuda())
    ref = [path.process_chunk(*c) for c in chunks]
    got = list(path.process_stream(iter(chunks)))
    assert len(got) == 3
    for a, b in zip(ref, got):
        assert np.array_equal(a["team"], b["team"]) and np.array_equal(a["count"], b["count"])
        assert len(a["team"]) == 12


def test_sliced_path_cuda_graph_replay_equals_eager(ctx):
    """SlicedPuckPath.process_chunk_device(graph=True): the whole chunk (K1b, per-shape-class YOLOv8n forwards with
    the K5 glue, K2a, gather, K2b) captured once and replayed gives what the eager launches give, also on new frames."""
    from hvb.pipeline import SlicedPuckPath
    from hvb.synth import rink_frame
    rng = np.random.default_rng(11)
    path = SlicedPuckPath("cuda:0", "n", 1, conf=5e-3, seed=2)            # random-init: low conf so that boxes come out
    f1 = torch.from_numpy(np.stack([rink_frame(rng, 720, 1280, 8)[0] for _ in range(2)])).cuda()
    f2 = torch.from_numpy(np.stack([rink_frame(rng, 720, 1280, 8)[0] for _ in range(2)])).cuda()
    eager1 = [t.clone() for t in path.process_chunk_device(f1)]
    eager2 = [t.clone() for t in path.process_chunk_device(f2)]
    before = ctx.launch_count()
    g1 = [t.clone() for t in path.process_chunk_device(f1, graph=True)]   # captures
    captured = ctx.launch_count() - before
    g2 = [t.clone() for t in path.process_chunk_device(f2, graph=True)]   # replays with other frames
    assert ctx.launch_count() - before == captured                         # the replay issued no eager libhvb launch
    for e, g in ((eager1, g1), (eager2, g2)):
        tot = int(e[4][-1])
        assert tot > 0 and int(g[4][-1]) == tot
        assert torch.equal(e[4], g[4]) and torch.equal(e[5], g[5])
        for a, b in zip(e[:4], g[:4]):
            assert torch.equal(a[:tot], b[:tot])


def test_sliced_stream_equals_run_batch(ctx):
    """SlicedPuckPath.process_stream (pinned H2D on a side stream, graph replay, async D2H) == run_batch per chunk."""
    from hvb.pipeline import SlicedPuckPath
    from hvb.synth import rink_frame
    rng = np.random.default_rng(13)
    path = SlicedPuckPath("cuda:0", "n", 1, conf=5e-3, seed=2)
    chunks = [np.stack([rink_frame(rng, 720, 1280, 8)[0] for _ in range(2)]) for _ in range(3)]
    want = [path.process_chunk(c) for c in chunks]
    for graph in (False, True):
        got = list(path.process_stream(iter(chunks), graph=graph))
        assert len(got) == len(want) == 3
        for g, w in zip(got, want):
            assert len(g) == len(w) == 2
            for a, b in zip(g, w):
                assert len(a) == len(b) and np.array_equal(a.xyxy, b.xyxy) and np.array_equal(a.confidence, b.confidence)
                assert np.array_equal(a.class_id, b.class_id)
