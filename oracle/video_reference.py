"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's two main-loop drivers.

Follows hockey/main.py: ``detect_players`` (:177-195), ``initialize_team_classifier`` (:197-257), ``process_frame``
(:259-313), ``process_video`` (:315-322), ``_get_crops`` / ``_get_positions`` / ``_create_color_lookup`` /
``_create_labels`` (:324-358), with ``TeamClassifier``'s routing to the hybrid classifier (common/team.py:134-271)
reduced to the path the default flags take when only the hybrid classifier is importable.

Pieces: the restated ultralytics predict path (oracle/ultralytics_restated.py, parity unpinned — library absent),
the restated supervision ByteTrack (oracle/bytetrack_restated.py, parity unpinned), and the team classifier
restatement that IS pinned against the real reference code (oracle/team_reference.py, tests/golden).  The YOLO
forward is injected as ``heads_fn(x_nchw float32) -> [3 raw head tensors]`` so the oracle and the GPU path can be fed
identical head tensors.  Never imported by the product.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import supervision_restated as svr
from . import ultralytics_restated as ur
from .bytetrack_restated import ByteTrack
from .team_reference import HybridReference

PLAYER_CLASS_ID, GOALKEEPER_CLASS_ID = 0, 1


@dataclass
class RefDetections:
    xyxy: np.ndarray
    confidence: np.ndarray
    class_id: np.ndarray
    tracker_id: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.xyxy)

    def take(self, idx):
        return RefDetections(self.xyxy[idx], self.confidence[idx], self.class_id[idx],
                             None if self.tracker_id is None else self.tracker_id[idx])

    @staticmethod
    def merge(items: Sequence["RefDetections"]) -> "RefDetections":
        items = [d for d in items if len(d)]
        if not items:
            return RefDetections(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), np.zeros((0,), int), np.zeros((0,), int))
        return RefDetections(np.vstack([d.xyxy for d in items]), np.concatenate([d.confidence for d in items]),
                             np.concatenate([d.class_id for d in items]), np.concatenate([d.tracker_id for d in items]))


@dataclass
class RefFrameResult:
    detections: RefDetections
    player_team_ids: np.ndarray
    goalie_team_ids: np.ndarray
    color_lookup: np.ndarray
    labels: List[str]


class VideoReference:
    def __init__(self, heads_fn: Callable[[torch.Tensor], List[torch.Tensor]], nc: int, trunk: torch.nn.Module,
                 imgsz: int = 1280, conf: float = 0.4, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: int = 30, minimum_consecutive_frames: int = 2,
                 initialization_stride: int = 10, max_initialization_frames: int = 20, min_players_for_selection: int = 6):
        self.heads_fn, self.nc, self.imgsz, self.conf = heads_fn, nc, imgsz, conf
        self.team = HybridReference(trunk)
        self.team_names = {0: "Team 0", 1: "Team 1"}
        self.frame_rate, self.track_activation_threshold = frame_rate, track_activation_threshold
        self.stride, self.max_init, self.min_players = initialization_stride, max_initialization_frames, min_players_for_selection
        self.tracker = ByteTrack(track_activation_threshold, lost_track_buffer, minimum_matching_threshold, frame_rate,
                                 minimum_consecutive_frames)

    # main.py:177-195
    def detect_players(self, frame: np.ndarray) -> RefDetections:
        lb = ur.letterbox(frame, self.imgsz, auto=True)
        x = torch.from_numpy(ur.preprocess([lb]))
        xyxy, conf, cls = ur.predict_from_head(self.heads_fn(x), self.nc, tuple(x.shape[2:]), [frame.shape[:2]], self.conf)[0]
        d = RefDetections(np.asarray(xyxy, np.float32).reshape(-1, 4), np.asarray(conf, np.float32), np.asarray(cls).astype(int))
        keep = ((d.class_id == PLAYER_CLASS_ID) | (d.class_id == GOALKEEPER_CLASS_ID)) & (d.confidence > self.conf)
        return d.take(keep)

    @staticmethod
    def _track(tracker: ByteTrack, d: RefDetections) -> RefDetections:
        keep, ids = tracker.update_with_detections(d.xyxy, d.confidence)
        out = d.take(keep)
        out.tracker_id = ids
        return out

    # main.py:197-257
    def initialize_team_classifier(self, frames: Iterable[np.ndarray]) -> None:
        crops, positions = [], []
        temp = ByteTrack(track_activation_threshold=self.track_activation_threshold, minimum_consecutive_frames=1,
                         frame_rate=self.frame_rate)
        sampled = (f for k, f in enumerate(frames) if k % self.stride == 0)
        for i, frame in enumerate(sampled):
            if i > self.max_init:
                break
            det = self.detect_players(frame)
            players = det.take(det.class_id == PLAYER_CLASS_ID)
            self._track(temp, players)
            crops.extend(svr.crop_image(frame, b) for b in players.xyxy)
            positions.extend(((b[0] + b[2]) / 2, (b[1] + b[3]) / 2) for b in players.xyxy)
        self.team.fit(crops)                     # TeamClassifier.fit does not forward positions to the hybrid classifier (team.py:193)
        self.n_fit_crops = len(crops)

    # main.py:259-313
    def process_frame(self, frame: np.ndarray) -> RefFrameResult:
        tracked = self._track(self.tracker, self.detect_players(frame))
        players = tracked.take(tracked.class_id == PLAYER_CLASS_ID)
        goalies = tracked.take(tracked.class_id == GOALKEEPER_CLASS_ID)
        team_ids = np.array([])
        if len(players):
            team_ids = self.team.predict([svr.crop_image(frame, b) for b in players.xyxy], tracker_ids=players.tracker_id)
        goalie_ids = np.array([2] * len(goalies), dtype=np.int32)
        merged = RefDetections.merge([players, goalies])
        lookup = np.concatenate([team_ids, goalie_ids]).astype(np.int32) if len(team_ids) else goalie_ids.astype(np.int32)
        labels = []
        for i, cid in enumerate(merged.class_id):
            if cid == PLAYER_CLASS_ID and i < len(team_ids):
                labels.append(self.team_names.get(int(team_ids[i]), f"Team {int(team_ids[i])}"))
            elif cid == GOALKEEPER_CLASS_ID:
                labels.append("Goalie")
            else:
                labels.append("Player")
        return RefFrameResult(merged, team_ids, goalie_ids, lookup, labels)

    # main.py:315-322
    def process_video(self, frames: Sequence[np.ndarray]):
        self.initialize_team_classifier(frames)
        return [self.process_frame(f) for f in frames]
