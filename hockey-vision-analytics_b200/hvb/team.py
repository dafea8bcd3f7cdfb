"""TeamClassifier — B200 drop-in for the router class the reference's main loop constructs
(hockey/common/team.py:37-330; used at hockey/main.py:158, 251-256, 277-281).

Same constructor flags, ``fit(crops, positions, frame, detections)``, ``predict(crops, tracker_ids,
positions)``, ``classify_jersey``, ``set_team_names`` / ``get_team_name`` / ``get_segmentation_masks``
and the same failure cascade: an exception inside the hybrid classifier permanently downgrades to
the simple HSV rule (team.py:192-198, 264-271).  The segmentation (GrabCut), interactive (GUI) and
robust (SigLIP + HDBSCAN) classifiers are outside the hot path and are reported as
unavailable, exactly as the reference behaves when their imports fail; the hybrid classifier is
therefore what the default flags route to.

Opt-in: ``use_segmentation="rectangle"`` routes to the GrabCut-free segmentation classifier
(hvb/team_segmentation.py, K3c) with the reference's cascade for that branch (team.py:141-154, 227-238):
a failure downgrades to the interactive classifier, which is unavailable, so to the simple rule.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional

import numpy as np

from . import _ffi
from .hybrid import HybridTeamClassifier
from .runtime import get_context
from .synth import pack_crops

HYBRID_AVAILABLE = True
ROBUST_AVAILABLE = False
INTERACTIVE_AVAILABLE = False
SEGMENTATION_AVAILABLE = False


class TeamClassifier:
    def __init__(self, device: str = "cuda:0", batch_size: int = 32, use_hybrid: bool = True, use_robust: bool = True,
                 use_interactive: bool = True, use_segmentation: bool = True, trunk=None):
        self.device = device
        self.batch_size = batch_size
        self.ctx = get_context(device)
        self.use_segmentation = bool(use_segmentation == "rectangle" or (use_segmentation and SEGMENTATION_AVAILABLE))
        self.use_interactive = use_interactive and INTERACTIVE_AVAILABLE and not self.use_segmentation
        self.use_robust = use_robust and ROBUST_AVAILABLE and not self.use_interactive and not self.use_segmentation
        self.use_hybrid = (use_hybrid and HYBRID_AVAILABLE and not self.use_robust and not self.use_interactive
                           and not self.use_segmentation)
        self.team_names = {0: "Team 0", 1: "Team 1"}
        self._trunk = trunk
        # state of the simple rule (also the landing spot of the failure cascade)
        self.player_history: Dict[int, List[int]] = defaultdict(list)
        self.history_window = 10
        self.team_assignments = {0: "away", 1: "home"}
        if self.use_segmentation:
            from .team_segmentation import SegmentationTeamClassifier
            self.segmentation_classifier = SegmentationTeamClassifier(device=device, visualize_segmentation=True)
        elif self.use_hybrid:
            self.hybrid_classifier = HybridTeamClassifier(device=device, trunk=trunk)

    # ------------------------------------------------------------------ simple HSV rule
    def extract_jersey_region(self, crop: np.ndarray) -> np.ndarray:
        h, w = crop.shape[:2]
        if h < 30 or w < 20:
            return crop
        region = crop[int(h * 0.25):int(h * 0.75), int(w * 0.3):int(w * 0.7)]
        return crop if region.size == 0 else region

    def _simple_stats(self, crops: List[np.ndarray]) -> np.ndarray:
        """COLOR_RAW rows over the simple-rule ROI (K3a kernel, ROI_SIMPLE)."""
        buf, desc = pack_crops(crops)
        cd = np.zeros((len(crops),), _ffi.CROP_DESC)
        cd["offset"], cd["pitch"], cd["h"], cd["w"] = desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3]
        _, raw = self.ctx.color_features_host(buf, cd, _ffi.ROI_SIMPLE, want_raw=True)
        if (raw["n"] == 0).any():
            raise ValueError("empty crop passed to classify_jersey")
        return raw

    @staticmethod
    def _rule(raw_row) -> tuple:
        n = float(raw_row["n"])
        avg_brightness = float(raw_row["sums"][2]) / n
        avg_saturation = float(raw_row["sums"][1]) / n
        white_ratio = float(raw_row["counts"][2]) / n
        if white_ratio > 0.3 or (avg_brightness > 180 and avg_saturation < 50):
            return 0, min(white_ratio * 2, 1.0)
        return 1, min(avg_saturation / 150, 1.0)

    def classify_jersey(self, crop: np.ndarray) -> tuple:
        return self._rule(self._simple_stats([crop])[0])

    # ------------------------------------------------------------------ fit / predict routing
    def fit(self, crops: List[np.ndarray], positions: Optional[List[tuple]] = None, frame: Optional[np.ndarray] = None,
            detections=None) -> None:
        if self.use_segmentation:
            try:
                self.segmentation_classifier.fit(crops, positions=positions)
            except Exception as e:                          # noqa: BLE001 - mirrors the reference cascade
                print(f"Segmentation classifier failed: {e}")
                print("Falling back to interactive classifier")
                self.use_segmentation = False              # INTERACTIVE_AVAILABLE is False -> simple rule (team.py:147-154)
                self._simple_fit(crops)
        elif self.use_hybrid:
            try:
                self.hybrid_classifier.fit(crops)          # positions are NOT forwarded (team.py:193)
            except Exception as e:                          # noqa: BLE001 - mirrors the reference cascade
                print(f"Hybrid classifier failed: {e}")
                print("Falling back to simple classifier")
                self.use_hybrid = False
                self._simple_fit(crops)
        else:
            self._simple_fit(crops)

    def _simple_fit(self, crops: List[np.ndarray]) -> None:
        sample = list(crops[:100])
        if not sample:
            return
        teams = [self._rule(r)[0] for r in self._simple_stats(sample)]
        white = sum(1 for t in teams if t == 0)
        print(f"Sample distribution - White jerseys: {white}, Colored jerseys: {len(teams) - white}")

    def predict(self, crops: List[np.ndarray], tracker_ids: Optional[np.ndarray] = None,
                positions: Optional[List[tuple]] = None) -> np.ndarray:
        if not len(crops):
            return np.array([])
        if self.use_segmentation:
            try:
                return self.segmentation_classifier.predict(crops, tracker_ids, positions)
            except Exception as e:                          # noqa: BLE001
                print(f"Segmentation prediction failed: {e}")
                print("Falling back to interactive classifier")
                self.use_segmentation = False
        if self.use_hybrid:
            try:
                return self.hybrid_classifier.predict(crops, tracker_ids)
            except Exception as e:                          # noqa: BLE001
                print(f"Hybrid prediction failed: {e}")
                print("Falling back to simple classifier")
                self.use_hybrid = False
        raw = self._simple_stats(list(crops))
        predictions = []
        for i in range(len(crops)):
            team, _ = self._rule(raw[i])
            if tracker_ids is not None and i < len(tracker_ids) and tracker_ids[i] is not None:
                tid = int(tracker_ids[i])
                hist = self.player_history[tid]
                hist.append(team)
                if len(hist) > self.history_window:
                    self.player_history[tid] = hist = hist[-self.history_window:]
                if len(hist) >= 3:
                    team = np.argmax(np.bincount(hist))
            predictions.append(team)
        return np.array(predictions)

    def predict_from_frame(self, frames_dev, xyxy, frame_idx=None, tracker_ids: Optional[np.ndarray] = None,
                           host_frames: Optional[np.ndarray] = None) -> np.ndarray:
        """predict() for crops that are boxes into device-resident frames (no host crops): same routing and the same
        failure cascade.  `host_frames` (the same frames on the host) is only needed if the cascade lands on the simple
        rule, which packs host crops."""
        if xyxy.shape[0] == 0:
            return np.array([])
        if self.use_segmentation:
            try:
                boxes_h = xyxy.cpu().numpy() if hasattr(xyxy, "cpu") else np.asarray(xyxy)
                fi_h = frame_idx.cpu().numpy() if hasattr(frame_idx, "cpu") else frame_idx
                return self.segmentation_classifier.predict_from_frames(frames_dev, boxes_h, fi_h, tracker_ids)
            except Exception as e:                          # noqa: BLE001
                print(f"Segmentation prediction failed: {e}")
                print("Falling back to interactive classifier")
                self.use_segmentation = False
        if self.use_hybrid:
            try:
                return self.hybrid_classifier.predict_from_frame(frames_dev, xyxy, frame_idx, tracker_ids)
            except Exception as e:                          # noqa: BLE001
                print(f"Hybrid prediction failed: {e}")
                print("Falling back to simple classifier")
                self.use_hybrid = False
        if host_frames is None:
            host_frames = frames_dev.cpu().numpy()
        from .detections import crop_image
        fi = frame_idx.cpu().numpy() if frame_idx is not None else np.zeros(xyxy.shape[0], np.int64)
        boxes = xyxy.cpu().numpy() if hasattr(xyxy, "cpu") else np.asarray(xyxy)
        frames = host_frames if isinstance(host_frames, (list, tuple)) or host_frames.ndim == 4 else host_frames[None]
        return self.predict([crop_image(frames[int(f)], b) for f, b in zip(fi, boxes)], tracker_ids)

    def get_segmentation_masks(self, tracker_ids: List[int]):
        if self.use_segmentation and hasattr(self, "segmentation_classifier"):
            return self.segmentation_classifier.get_segmentation_masks(tracker_ids)
        return None

    def set_team_names(self, team_names: Dict[int, str]) -> None:
        self.team_names.update(team_names)

    def get_team_name(self, team_id: int) -> str:
        return self.team_names.get(team_id, f"Team {team_id}")
