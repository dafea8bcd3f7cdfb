"""Probe: where does the 4K sliced e2e step go?  H2D bandwidth of the pinned chunk, graph replay alone, the pipelined
process_stream with per-chunk wall-clock marks.  Run on a GPU box: python tools/probe_e2e4k.py"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hockey-vision-analytics_b200")]
import torch
from hvb.pipeline import SlicedPuckPath
from hvb.synth import rink_clip

dev = "cuda:0"
F4 = 16
f4 = rink_clip(7, F4, 2160, 3840, 12, 2.0)[0]
pinned = torch.from_numpy(f4).pin_memory()
out = {}
d = torch.empty(pinned.shape, dtype=torch.uint8, device=dev)
for tag, src in (("pinned", pinned), ("pageable", torch.from_numpy(f4))):
    for _ in range(2):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    out["h2d_GBps_" + tag] = 5 * pinned.numel() / (time.perf_counter() - t0) / 1e9
h = torch.empty(pinned.shape, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); out["d2h_GBps_pinned"] = pinned.numel() / (time.perf_counter() - t0) / 1e9

puck = SlicedPuckPath(dev, "n", 1, 0.4)
for _ in range(3):
    puck.process_chunk_device(d, graph=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    puck.process_chunk_device(d, graph=True)
torch.cuda.synchronize()
out["graph_replay_ms"] = 1e3 * (time.perf_counter() - t0) / 5
# replay + concurrent H2D on a side stream
s = torch.cuda.Stream()
d2 = torch.empty_like(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s):
        d2.copy_(pinned, non_blocking=True)
    puck.process_chunk_device(d, graph=True)
torch.cuda.synchronize()
out["graph_replay_with_concurrent_h2d_ms"] = 1e3 * (time.perf_counter() - t0) / 5

marks = []
list(puck.process_stream(pinned for _ in range(2)))
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in puck.process_stream(pinned for _ in range(8)):
    marks.append(round(1e3 * (time.perf_counter() - t0), 2))
out["process_stream_marks_ms"] = marks
print(json.dumps(out))

# ---- frame-at-a-time process_frame: phases (host wall clock with a device sync after each phase)
from hvb.pipeline import HotPath
from hvb.video import Config, VideoProcessor
from hvb.synth import PlantedOverlay
from hvb.detect import PLAYER_CLASS_ID
H, W = 1080, 1920
frames, boxes, _, _ = rink_clip(1000, 4, H, W, 12, 1.0)
cls = [np.array([0] * 11 + [1])] * 4
path = HotPath(dev, "m", 2, 1280, 0.4, seed=0)
det = path.detector
ov = PlantedOverlay.whole_frame(77, (H, W), 1280, boxes, cls, nc=2, dup=3)
det.head_hook = ov.to_device(dev, [[0]])
sk = np.concatenate([b[:11] for b in boxes]).astype(np.float32)
path.classifier.fit_features(path.classifier.features_from_frame(torch.from_numpy(frames).to(dev), torch.from_numpy(sk).to(dev),
                             torch.from_numpy(np.repeat(np.arange(4, dtype=np.int32), 11)).to(dev))[0], None, None, cluster=False)
det.cuda_graph = True
vp = VideoProcessor(device=dev, config=Config(), detector=det, team_classifier=path.classifier_router(), tracker="device")
for _ in range(5):
    vp.process_frame(frames[0])
ph = {k: 0.0 for k in ("upload", "detect", "track", "team", "finish")}
sync = torch.cuda.synchronize
for _ in range(20):
    t = [time.perf_counter()]
    fd = det.upload(frames[0]); sync(); t.append(time.perf_counter())
    d = det.detect_players(fd); sync(); t.append(time.perf_counter())
    tr = vp.tracker.update_with_detections(d); sync(); t.append(time.perf_counter())
    pl = tr[tr.class_id == PLAYER_CLASS_ID]; gl = tr[tr.class_id != PLAYER_CLASS_ID]
    ids = vp.team_classifier.predict_from_frame(fd, torch.from_numpy(np.ascontiguousarray(pl.xyxy, np.float32)), None,
                                                tracker_ids=pl.tracker_id, host_frames=[frames[0]]); sync(); t.append(time.perf_counter())
    vp._finish(pl, gl, ids); t.append(time.perf_counter())
    for k, a, b in zip(ph, t[:-1], t[1:]):
        ph[k] += 1e3 * (b - a) / 20
out2 = {"process_frame_phase_ms": {k: round(v, 3) for k, v in ph.items()}}
# team stage sub-phases
clf = path.classifier
xy = torch.from_numpy(np.ascontiguousarray(pl.xyxy, np.float32))
sub = {k: 0.0 for k in ("crops_from_boxes", "K3a", "K3b", "trunk", "rest")}
from hvb import _ffi
ctx = clf.ctx
for _ in range(20):
    t = [time.perf_counter()]
    cd = ctx.crops_from_boxes(xy.to(dev), None, H, W); sync(); t.append(time.perf_counter())
    f = ctx.color_features(fd, cd, len(xy), _ffi.ROI_HYBRID); sync(); t.append(time.perf_counter())
    x, valid = ctx.mnv3_preprocess(fd, cd, len(xy), _ffi.ROI_HYBRID, rows=16); sync(); t.append(time.perf_counter())
    deep = clf._trunk_forward(x); sync(); t.append(time.perf_counter())
    clf._predict_device(fd, cd, len(xy), None); sync(); t.append(time.perf_counter())
    for k, a, b in zip(sub, t[:-1], t[1:]):
        sub[k] += 1e3 * (b - a) / 20
out2["team_stage_ms (rest = the whole _predict_device again)"] = {k: round(v, 3) for k, v in sub.items()}
print(json.dumps(out2))
