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
from .tracker import ByteTrack, DeviceByteTrack


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
    def __init__(self, player_model: Optional[torch.nn.Module] = None, device: str = "cuda:0", config: Optional[Config] = None,
                 team_classifier: Optional[TeamClassifier] = None, trunk: Optional[torch.nn.Module] = None,
                 team_selector: Optional[Callable] = None, detector_kwargs: Optional[dict] = None,
                 detector: Optional[Detector] = None, tracker: str = "device"):
        """`tracker`: "device" = K7 (the whole of sv.ByteTrack in one kernel, fed from K2a's device outputs in the chunked
        path); "host" = hvb.tracker.ByteTrack (numpy bookkeeping, K4b cost matrices).  Same results.
        `detector`: reuse an existing Detector (its weights, plans and cuDNN autotuning) instead of building one."""
        self.config = config or Config()
        self.device = device
        if detector is None:
            kw = dict(fuse=True, channels_last=True, cuda_graph=None)   # frame-at-a-time calls are launch-bound: replay a graph when the forward is capturable
            kw.update(detector_kwargs or {})
            detector = Detector(player_model, device, imgsz=self.config.detection_imgsz, conf=self.config.detection_confidence,
                                class_names={PLAYER_CLASS_ID: "player", GOALKEEPER_CLASS_ID: "goalie"}, **kw)
        self.detector = detector
        self.team_classifier = team_classifier if team_classifier is not None else TeamClassifier(device=device, trunk=trunk)
        self.team_selector = team_selector
        if tracker not in ("device", "host"):
            raise ValueError("tracker must be 'device' or 'host'")
        self.tracker_backend = tracker
        c = self.config
        self.tracker = self._make_tracker(track_activation_threshold=c.track_activation_threshold, lost_track_buffer=c.lost_track_buffer,
                                          minimum_matching_threshold=c.minimum_matching_threshold, frame_rate=c.frame_rate,
                                          minimum_consecutive_frames=c.minimum_consecutive_frames)
        self._track_stream = None
        self._team_stream = None
        self._pinned_out = None

    def _make_tracker(self, **kw):
        if self.tracker_backend == "device":
            return DeviceByteTrack(device=self.device, **kw)
        return ByteTrack(device=self.device, **kw)

    # ------------------------------------------------------------------ main.py:177-195
    def detect_players(self, frame: np.ndarray) -> Detections:
        return self.detector.detect_players(frame)

    # ------------------------------------------------------------------ main.py:197-257
    def initialize_team_classifier(self, frames: Iterable[np.ndarray]) -> None:
        print("Initializing team classification...")
        c = self.config
        crops, positions = [], []
        first_frame, first_tracked = None, None
        temp_tracker = self._make_tracker(track_activation_threshold=c.track_activation_threshold, minimum_consecutive_frames=1,
                                          frame_rate=c.frame_rate)
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
        # the frame crosses PCIe once: the detector and the team stage both read the device copy (crops are boxes into it,
        # hvb_crops_from_boxes == sv.crop_image's rounding and slicing), instead of packing host crops and uploading them again
        frame_dev = self.detector.upload(frame)
        detections = self.detector.detect_players(frame_dev)
        tracked = self.tracker.update_with_detections(detections)
        players = tracked[tracked.class_id == PLAYER_CLASS_ID]
        goalies = tracked[tracked.class_id == GOALKEEPER_CLASS_ID]
        player_team_ids = np.array([])
        if len(players) > 0:
            player_team_ids = self.team_classifier.predict_from_frame(
                frame_dev, torch.from_numpy(np.ascontiguousarray(players.xyxy, np.float32)), None,
                tracker_ids=players.tracker_id, host_frames=[frame])
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
        a side stream; the GPU detects and tracks chunk i (results leave through an asynchronous D2H copy + an event);
        the main thread runs the team stage of chunk i-1.  Results are yielded in frame order."""
        import queue
        import threading
        if initialize:
            self.initialize_team_classifier(frames)
        det = self.detector
        dev = det.device
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

        def source():
            while True:
                item = q.get()
                if isinstance(item, BaseException):
                    raise item
                if item is None:
                    return
                yield item

        yield from self.process_chunks(source())

    def process_chunks(self, chunks: Iterable) -> Iterator[FrameResult]:
        """The chunk pipeline behind process_video_chunked, for chunks that are (or are being made) resident on the device.

        `chunks` yields either a device tensor uint8[n,H,W,3] or a tuple (host_block | None, frames_dev, ready_event | None)
        (`ready_event`: the H2D copy of that chunk, recorded on another stream).  Per chunk, stream-ordered and without
        a host round trip: K1a -> YOLO -> K2a on the main stream, then (device tracker) K7 on a side stream stepping the
        chunk's frames in order straight from K2a's outputs, then one asynchronous D2H of detections + tracker ids.  The host
        touches chunk i-1 (player boxes -> team stage K3a/K3b/MobileNetV3/K4a -> rule + temporal vote, FrameResults) while
        the GPU works on chunk i."""
        det = self.detector
        dev = det.device
        main = torch.cuda.current_stream(dev)
        if self._track_stream is None:
            self._track_stream = torch.cuda.Stream(device=dev, priority=-1)
        side = self._track_stream
        if self._team_stream is None:
            # the team stage of chunk i-1 is issued AFTER chunk i's detection has been queued on the main stream; on its
            # own high-priority stream it runs beside that detection instead of behind it (the host waits for its result)
            self._team_stream = torch.cuda.Stream(device=dev, priority=-1)
        team_stream = self._team_stream
        if self._pinned_out is None:                                    # page-locked result buffers live as long as the processor:
            self._pinned_out = [dict(), dict(), dict()]                 # cudaHostAlloc per call cost ~10 ms per 384-frame clip (cProfile, run r02zb)
        pinned = self._pinned_out
        device_tracker = self.tracker_backend == "device"
        conf_thr = float(self.config.detection_confidence)
        cmask = (1 << PLAYER_CLASS_ID) | (1 << GOALKEEPER_CLASS_ID)

        def track(ci, dets, seq):
            """K7 on the tracker stream + the asynchronous D2H of everything the host reads for this chunk."""
            xyxy, cf, cl, cnt = dets
            outs = [("xyxy", xyxy), ("conf", cf), ("cls", cl), ("count", cnt)]
            stream = main
            if device_tracker:
                side.wait_stream(main)
                stream = side
                with torch.cuda.stream(side):
                    row, tid, tcnt = self.tracker.update_chunk_device(xyxy, cf, cl, cnt, min_conf=conf_thr, class_mask=cmask, seq=seq)
                outs += [("row", row), ("tid", tid), ("tcount", tcnt)]
                for t in dets:
                    t.record_stream(side)
            host = pinned[ci % 3]
            with torch.cuda.stream(stream):
                for k, t in outs:
                    if k not in host or host[k].shape != t.shape:
                        host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    host[k].copy_(t, non_blocking=True)
                done = torch.cuda.Event()
                done.record(stream)
            return host, done

        def launch(ci, item):
            block, frames_dev, ready = item if isinstance(item, tuple) else (None, item, None)
            if ready is not None:
                main.wait_event(ready)
                frames_dev.record_stream(main)
            xyxy, cf, cl, cnt, state = det.detect_device(frames_dev)
            if det.cuda_graph:                                          # static graph outputs: the next replay overwrites them
                xyxy, cf, cl, cnt = xyxy.clone(), cf.clone(), cl.clone(), cnt.clone()
            seq = self.tracker.next_seq() if device_tracker else 0
            host, done = track(ci, (xyxy, cf, cl, cnt), seq)
            return dict(ci=ci, block=block, frames=frames_dev, host=host, done=done, dets=(xyxy, cf, cl, cnt), state=state,
                        ready=ready, seq=seq)

        def finish(p, nxt):
            block, frames_dev, host, ready = p["block"], p["frames"], p["host"], p["ready"]
            xyxy, cf, cl, cnt = p["dets"]
            p["done"].synchronize()
            cnt_h = host["count"].numpy().copy()
            redo = False
            if (cnt_h < 0).any():                                       # > 1024 candidates in a frame: K2a's large tier for those frames
                state = p["state"]
                if det.cuda_graph:                                      # the graph's head tensors now hold a later chunk: redo eagerly
                    if det.head_hook is not None:
                        raise RuntimeError("candidate overflow with a head_hook under CUDA graphs is not supported")
                    xyxy, cf, cl, cnt, state = det.detect_device(frames_dev, graph=False)
                    cnt_h = cnt.cpu().numpy()
                cnt_h = det._retry_overflow(xyxy, cf, cl, cnt, state, cnt_h)
                redo = True
            if device_tracker and (redo or (host["tcount"].numpy() == -2).any()):
                # K7 rejected this chunk (its own overflow, or it was queued behind a rejected chunk) and left the clip
                # untouched: step it now, then resubmit the chunk already queued behind it so the pipeline stays in order
                host, done = track(p["ci"], (xyxy, cf, cl, cnt), p["seq"])
                done.synchronize()
                if nxt is not None:
                    nxt["host"], nxt["done"] = track(nxt["ci"], nxt["dets"], nxt["seq"])
            elif redo:
                host = dict(xyxy=xyxy.cpu(), conf=cf.cpu(), cls=cl.cpu())
            xyxy_h, cf_h, cl_h = host["xyxy"].numpy(), host["conf"].numpy(), host["cls"].numpy()
            n = len(cnt_h)
            if device_tracker:
                tc = host["tcount"].numpy()
                if (tc < 0).any():
                    from . import _ffi
                    raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "ByteTrack capacity exceeded (256 live tracks per clip, 320 detections per frame)")
                row_h, tid_h = host["row"].numpy(), host["tid"].numpy()
                fi, pos = np.nonzero(np.arange(row_h.shape[1])[None, :] < tc[:, None])       # frame-major, detection order
                rows = row_h[fi, pos]
                t_xyxy, t_conf, t_cls, t_tid = xyxy_h[fi, rows], cf_h[fi, rows], cl_h[fi, rows].astype(int), tid_h[fi, pos].astype(int)
                seg = np.concatenate([[0], np.cumsum(tc)])
            else:
                parts = []
                for i, k in enumerate(cnt_h):                           # host ByteTrack: strictly sequential per frame
                    d = det._to_detections(xyxy_h[i, :k], cf_h[i, :k], cl_h[i, :k])
                    d = d[((d.class_id == PLAYER_CLASS_ID) | (d.class_id == GOALKEEPER_CLASS_ID)) & (d.confidence > conf_thr)]
                    parts.append(self.tracker.update_with_detections(d))
                tc = np.array([len(d) for d in parts])
                fi = np.repeat(np.arange(n), tc)
                seg = np.concatenate([[0], np.cumsum(tc)])
                cat = lambda f, dt: np.concatenate([np.asarray(f(d)) for d in parts]).astype(dt) if len(fi) else np.zeros((0,), dt)
                t_xyxy = np.concatenate([np.asarray(d.xyxy, np.float32).reshape(-1, 4) for d in parts]) if len(fi) else np.zeros((0, 4), np.float32)
                t_conf, t_cls, t_tid = cat(lambda d: d.confidence, np.float32), cat(lambda d: d.class_id, int), cat(lambda d: d.tracker_id, int)
            is_player = t_cls == PLAYER_CLASS_ID
            team_ids = np.array([])
            if is_player.any():                                         # one feature pass for every tracked player of the chunk
                if ready is not None:
                    team_stream.wait_event(ready)                       # the chunk's H2D copy
                frames_dev.record_stream(team_stream)
                with torch.cuda.stream(team_stream):
                    team_ids = self.team_classifier.predict_from_frame(
                        frames_dev, torch.from_numpy(np.ascontiguousarray(t_xyxy[is_player], np.float32)),
                        torch.from_numpy(fi[is_player].astype(np.int32)).to(dev), tracker_ids=t_tid[is_player],
                        host_frames=block)                              # host frames only for the fallback cascade
            names = det.class_names
            out, tpos = [], 0
            for i in range(n):
                lo, hi = int(seg[i]), int(seg[i + 1])
                pl = np.nonzero(is_player[lo:hi])[0] + lo
                gl = np.nonzero(t_cls[lo:hi] == GOALKEEPER_CLASS_ID)[0] + lo
                order = np.concatenate([pl, gl])                        # Detections.merge([players, goalies])
                cls_i = t_cls[order]
                if len(order):
                    merged = Detections(xyxy=t_xyxy[order], confidence=t_conf[order], class_id=cls_i, tracker_id=t_tid[order],
                                        data={"class_name": np.array([names.get(int(c), str(int(c))) for c in cls_i], dtype=object)})
                else:
                    merged = Detections.empty()
                ptid = team_ids[tpos:tpos + len(pl)] if len(pl) else np.array([])
                tpos += len(pl)
                gtid = np.array([2] * len(gl), dtype=np.int32)
                out.append(FrameResult(merged, ptid, gtid, self._create_color_lookup(ptid, gtid), self._create_labels(merged, ptid)))
            return out

        pending = None
        for ci, item in enumerate(chunks):
            nxt = launch(ci, item)                                      # chunk ci is queued on the GPU ...
            if pending is not None:
                yield from finish(pending, nxt)                         # ... while the host finishes chunk ci-1
            pending = nxt
        if pending is not None:
            yield from finish(pending, None)

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
