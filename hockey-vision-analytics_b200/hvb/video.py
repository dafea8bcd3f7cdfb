"""VideoProcessor — the two drivers of the reference's main loop, on the B200 path.

Mirrors ``hockey/main.py``: ``Config`` (:18-59, hot-path fields), ``VideoProcessor.detect_players`` (:177-195),
``initialize_team_classifier`` (:197-257: frames sampled with stride 10, at most 21 of them, a temporary ByteTrack,
crops + positions of the player class, then ``TeamClassifier.fit``), ``process_frame`` (:259-313: detect -> ByteTrack
-> player / goalie split -> crops -> ``TeamClassifier.predict(crops, tracker_ids, positions)`` -> goalies = team 2 ->
merge -> colour lookup + labels) and ``process_video`` (:315-322).  Same call order, same masks, same sampling.

What is not here: annotation, display, video sinks, the interactive team selector GUI and the rink-keypoint branch
(presentation).  ``process_frame`` therefore returns the data the reference hands to its annotator
(``all_detections, labels, color_lookup`` — main.py:287-311) as a ``FrameResult`` instead of an annotated image, and
``initialize_team_classifier`` takes the clip's frames (any iterable) where the reference takes a path and opens it
with ``sv.get_video_frames_generator``; a ``team_selector`` callable can be injected, the default is the
"selection cancelled" branch.

``process_video_chunked`` is the fast path for clips: detection runs on a chunk of frames per launch (K1a -> YOLO ->
K2a), ByteTrack stays sequential per frame on the host, and the team features of ALL tracked players of the chunk
come from one K3a/K3b/MobileNetV3/K4a pass over the device-resident frames; the temporal vote is then applied in
frame order, so the results are the ones frame-at-a-time processing gives.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .detect import Detector, GOALKEEPER_CLASS_ID, PLAYER_CLASS_ID
from .detections import Detections, crop_image
from .team import TeamClassifier
from .tracker import ByteTrack


@dataclass
class Config:
    """hockey/main.py:18-59 — the fields the hot path reads, with the reference's defaults."""
    detection_imgsz: int = 1280
    detection_confidence: float = 0.4
    track_activation_threshold: float = 0.25
    lost_track_buffer: int = 30
    minimum_matching_threshold: float = 0.8
    frame_rate: int = 30
    minimum_consecutive_frames: int = 2
    initialization_stride: int = 10
    max_initialization_frames: int = 20
    min_players_for_selection: int = 6


@dataclass
class FrameResult:
    """What process_frame hands to the annotator in the reference (main.py:287-311)."""
    detections: Detections                    # merge([player_detections, goalie_detections]), tracker_id set
    player_team_ids: np.ndarray               # one per player detection (first len(players) rows of `detections`)
    goalie_team_ids: np.ndarray               # all 2
    color_lookup: np.ndarray                  # int32, same order as `detections`
    labels: List[str] = field(default_factory=list)


class VideoProcessor:
    def __init__(self, player_model: torch.nn.Module, device: str = "cuda:0", config: Optional[Config] = None,
                 team_classifier: Optional[TeamClassifier] = None, trunk: Optional[torch.nn.Module] = None,
                 team_selector: Optional[Callable] = None, detector_kwargs: Optional[dict] = None):
        self.config = config or Config()
        self.device = device
        kw = dict(fuse=True, channels_last=True, cuda_graph=None)       # frame-at-a-time calls are launch-bound: replay a graph when the forward is capturable
        kw.update(detector_kwargs or {})
        self.detector = Detector(player_model, device, imgsz=self.config.detection_imgsz, conf=self.config.detection_confidence,
                                 class_names={PLAYER_CLASS_ID: "player", GOALKEEPER_CLASS_ID: "goalie"}, **kw)
        self.team_classifier = team_classifier if team_classifier is not None else TeamClassifier(device=device, trunk=trunk)
        self.team_selector = team_selector
        c = self.config
        self.tracker = ByteTrack(track_activation_threshold=c.track_activation_threshold, lost_track_buffer=c.lost_track_buffer,
                                 minimum_matching_threshold=c.minimum_matching_threshold, frame_rate=c.frame_rate,
                                 minimum_consecutive_frames=c.minimum_consecutive_frames, device=device)

    # ------------------------------------------------------------------ main.py:177-195
    def detect_players(self, frame: np.ndarray) -> Detections:
        return self.detector.detect_players(frame)

    # ------------------------------------------------------------------ main.py:197-257
    def initialize_team_classifier(self, frames: Iterable[np.ndarray]) -> None:
        print("Initializing team classification...")
        c = self.config
        crops, positions = [], []
        first_frame, first_tracked = None, None
        temp_tracker = ByteTrack(track_activation_threshold=c.track_activation_threshold, minimum_consecutive_frames=1,
                                 frame_rate=c.frame_rate, device=self.device)
        sampled = (f for k, f in enumerate(frames) if k % c.initialization_stride == 0)    # get_video_frames_generator(stride=)
        for i, frame in enumerate(sampled):
            if i > c.max_initialization_frames:
                break
            detections = self.detect_players(frame)
            player_detections = detections[detections.class_id == PLAYER_CLASS_ID]
            tracked = temp_tracker.update_with_detections(player_detections)
            if first_frame is None and len(tracked) >= c.min_players_for_selection:
                first_frame, first_tracked = frame, tracked
            crops.extend(self._get_crops(frame, player_detections))
            positions.extend(self._get_positions(player_detections))
        selection = None
        if first_frame is not None and first_tracked is not None and self.team_selector is not None:
            selection = self.team_selector(first_frame, first_tracked)
        if selection:
            self.team_classifier.set_team_names(selection.team_names)
            print(f"Teams set: {selection.team_names[0]} vs {selection.team_names[1]}")
        else:
            print("Team selection cancelled, using default team names")
        self.team_classifier.fit(crops, positions=positions, frame=first_frame, detections=first_tracked)
        print("Classifier fitted.")

    # ------------------------------------------------------------------ main.py:259-313
    def process_frame(self, frame: np.ndarray) -> FrameResult:
        detections = self.detect_players(frame)
        tracked = self.tracker.update_with_detections(detections)
        players = tracked[tracked.class_id == PLAYER_CLASS_ID]
        goalies = tracked[tracked.class_id == GOALKEEPER_CLASS_ID]
        player_team_ids = np.array([])
        if len(players) > 0:
            player_team_ids = self.team_classifier.predict(self._get_crops(frame, players), tracker_ids=players.tracker_id,
                                                           positions=self._get_positions(players))
        return self._finish(players, goalies, player_team_ids)

    def _finish(self, players: Detections, goalies: Detections, player_team_ids: np.ndarray) -> FrameResult:
        goalie_team_ids = np.array([2] * len(goalies), dtype=np.int32)
        merged = Detections.merge([players, goalies])
        return FrameResult(merged, player_team_ids, goalie_team_ids, self._create_color_lookup(player_team_ids, goalie_team_ids),
                           self._create_labels(merged, player_team_ids))

    # ------------------------------------------------------------------ main.py:315-322
    def process_video(self, frames: Sequence[np.ndarray]) -> Iterator[FrameResult]:
        self.initialize_team_classifier(frames)
        for frame in frames:
            yield self.process_frame(frame)

    def process_video_chunked(self, frames: Sequence[np.ndarray], chunk: int = 32, initialize: bool = True) -> Iterator[FrameResult]:
        """Same results as process_video, GPU work batched per chunk of frames (see the module docstring).

        Three things run concurrently: a staging thread copies chunk i+1 into a pinned buffer and starts its H2D copy on
        a side stream; the GPU detects chunk i (its results leave through an asynchronous D2H copy + an event); the main
        thread runs ByteTrack and the team stage of chunk i-1.  Results are yielded in frame order."""
        import queue
        import threading
        if initialize:
            self.initialize_team_classifier(frames)
        det, conf = self.detector, self.config.detection_confidence
        dev = det.device
        main = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        starts = list(range(0, len(frames), chunk))
        q: "queue.Queue" = queue.Queue(maxsize=2)

        def stager():
            try:
                with torch.cuda.device(dev):
                    for lo in starts:
                        block = frames[lo:lo + chunk]                   # a list of frames: staged without an extra copy
                        with torch.cuda.stream(copy_stream):
                            fd = det.upload(block)
                            ev = torch.cuda.Event()
                            ev.record(copy_stream)
                        q.put((block, fd, ev))
                q.put(None)
            except BaseException as e:                                  # noqa: BLE001 - surfaced in the consumer
                q.put(e)

        threading.Thread(target=stager, daemon=True).start()
        pinned = [dict(), dict()]

        def launch(ci, item):
            block, frames_dev, ready = item
            main.wait_event(ready)
            frames_dev.record_stream(main)
            xyxy, cf, cl, cnt, _state = det.detect_device(frames_dev)
            host = pinned[ci & 1]
            for k, t in (("xyxy", xyxy), ("conf", cf), ("cls", cl), ("count", cnt)):
                if k not in host or host[k].shape != t.shape:
                    host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                host[k].copy_(t, non_blocking=True)                    # stream-ordered before the next graph replay
            done = torch.cuda.Event()
            done.record(main)
            return block, frames_dev, host, done

        def finish(p):
            block, frames_dev, host, done = p
            done.synchronize()
            cnt_h = host["count"].numpy().copy()
            xyxy_h, cf_h, cl_h = host["xyxy"].numpy(), host["conf"].numpy(), host["cls"].numpy()
            if (cnt_h < 0).any():                                       # > 1024 candidates in a frame: redo eagerly with the retry tier
                xyxy, cf, cl, cnt, state = det.detect_device(frames_dev, graph=False)
                cnt_h = det._retry_overflow(xyxy, cf, cl, cnt, state, cnt.cpu().numpy())
                xyxy_h, cf_h, cl_h = xyxy.cpu().numpy(), cf.cpu().numpy(), cl.cpu().numpy()
            per_frame, boxes, fidx, tids = [], [], [], []
            for i, k in enumerate(cnt_h):                             # ByteTrack: strictly sequential per frame
                d = det._to_detections(xyxy_h[i, :k], cf_h[i, :k], cl_h[i, :k])
                d = d[((d.class_id == PLAYER_CLASS_ID) | (d.class_id == GOALKEEPER_CLASS_ID)) & (d.confidence > conf)]
                tracked = self.tracker.update_with_detections(d)
                players = tracked[tracked.class_id == PLAYER_CLASS_ID]
                goalies = tracked[tracked.class_id == GOALKEEPER_CLASS_ID]
                per_frame.append((players, goalies))
                if len(players):
                    boxes.append(np.asarray(players.xyxy, np.float32)); fidx.append(np.full(len(players), i, np.int32))
                    tids.append(np.asarray(players.tracker_id))
            team_ids = np.array([])
            if boxes:                                                   # one feature pass for every tracked player of the chunk
                team_ids = self.team_classifier.predict_from_frame(
                    frames_dev, torch.from_numpy(np.concatenate(boxes)), torch.from_numpy(np.concatenate(fidx)).to(dev),
                    tracker_ids=np.concatenate(tids), host_frames=block)          # host frames only for the fallback cascade
            out, pos = [], 0
            for players, goalies in per_frame:
                n = len(players)
                out.append(self._finish(players, goalies, team_ids[pos:pos + n] if n else np.array([])))
                pos += n
            return out

        pending, ci = None, 0
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            nxt = launch(ci, item) if item is not None else None        # chunk ci is queued on the GPU ...
            ci += 1
            if pending is not None:
                yield from finish(pending)                              # ... while the host finishes chunk ci-1
            pending = nxt
            if item is None:
                break

    # ------------------------------------------------------------------ main.py:324-358
    @staticmethod
    def _get_crops(frame: np.ndarray, detections: Detections) -> List[np.ndarray]:
        return [crop_image(frame, xyxy) for xyxy in detections.xyxy]

    @staticmethod
    def _get_positions(detections: Detections) -> List[Tuple[float, float]]:
        return [((xyxy[0] + xyxy[2]) / 2, (xyxy[1] + xyxy[3]) / 2) for xyxy in detections.xyxy]

    @staticmethod
    def _create_color_lookup(player_team_ids: np.ndarray, goalie_team_ids: np.ndarray) -> np.ndarray:
        if len(player_team_ids) > 0:
            return np.concatenate([player_team_ids, goalie_team_ids]).astype(np.int32)
        return goalie_team_ids.astype(np.int32)

    def _create_labels(self, detections: Detections, player_team_ids: np.ndarray) -> List[str]:
        labels = []
        tracker_ids = detections.tracker_id if detections.tracker_id is not None else [None] * len(detections)
        for i, (_tid, class_id) in enumerate(zip(tracker_ids, detections.class_id)):
            if class_id == PLAYER_CLASS_ID and i < len(player_team_ids):
                labels.append(self.team_classifier.get_team_name(player_team_ids[i]))
            elif class_id == GOALKEEPER_CLASS_ID:
                labels.append("Goalie")
            else:
                labels.append("Player")
        return labels
