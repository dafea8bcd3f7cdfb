"""TEST INFRASTRUCTURE ONLY — CPU restatement of the GrabCut-free part of the reference's
SegmentationTeamClassifier (hockey/common/team_segmentation.py), SURVEY.md §8(f) rank 4.

Follows: the rectangle ``segment_player`` falls back to when GrabCut raises (:87-96), ``extract_jersey_colors``
(:98-148), ``classify_single_jersey`` (:150-165), ``fit`` (:167-228) and ``predict`` (:230-292) with
``segment_player`` fixed to that rectangle.  Uses the same real cv2 / numpy / sklearn routines as the reference.

Pinned: tests/golden/segmentation_reference.npz holds the outputs of the REAL reference class (imported from
/root/reference by tests/golden/make_golden_segmentation.py with ``cv2.grabCut`` patched to raise, which is the
reference's own route into the fallback rectangle); tests/test_oracle_segmentation.py checks this restatement
against the file and, in the build container, against a live import.  Never imported by the product.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np
from sklearn.cluster import KMeans


def fallback_mask(height: int, width: int) -> np.ndarray:
    """team_segmentation.py:87-96."""
    m = np.zeros((height, width), dtype=bool)
    m[int(height * 0.2):int(height * 0.6), int(width * 0.3):int(width * 0.7)] = True
    return m


def extract_jersey_colors(crop: np.ndarray, mask: np.ndarray) -> Dict[str, float]:
    """team_segmentation.py:98-148.  NOTE ``a_channel - 128`` is uint8 arithmetic: values below 128 wrap, so the
    "near neutral" test only passes for a, b in [128, 138)."""
    px = crop[mask]
    if len(px) < 100:
        return {"is_white": 0.5, "dominant_hue": 0, "saturation": 0, "brightness": 128}
    hsv = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2HSV).reshape(-1, 3)
    lab = cv2.cvtColor(px.reshape(-1, 1, 3), cv2.COLOR_BGR2LAB).reshape(-1, 3)
    white = (lab[:, 0] > 200) & (np.abs(lab[:, 1] - 128) < 10) & (np.abs(lab[:, 2] - 128) < 10)
    white_ratio = np.sum(white) / len(white)
    colored = hsv[~white]
    if len(colored) > 50:
        hist, _ = np.histogram(colored[:, 0], bins=18, range=(0, 180))
        dominant_hue = np.argmax(hist) * 10
        sat = np.mean(colored[:, 1])
    else:
        dominant_hue = 0
        sat = np.mean(hsv[:, 1])
    return {"is_white": white_ratio, "dominant_hue": dominant_hue, "saturation": sat, "brightness": np.mean(hsv[:, 2])}


def feature_row(f: Dict[str, float]) -> List[float]:
    return [f["is_white"], f["dominant_hue"], f["saturation"], f["brightness"]]


def classify_single_jersey(crop: np.ndarray) -> Tuple[int, float]:
    """team_segmentation.py:150-165."""
    f = extract_jersey_colors(crop, fallback_mask(*crop.shape[:2]))
    if f["is_white"] > 0.4:
        return 0, f["is_white"]
    return 1, min(f["saturation"] / 150, 1.0)


class SegmentationReference:
    """fit / predict of team_segmentation.py:167-292 with the rectangle mask."""

    def __init__(self):
        self.player_history: Dict[int, List[int]] = defaultdict(list)
        self.history_window = 10
        self.kmeans = None
        self.team_colors = None

    def fit(self, crops: List[np.ndarray]) -> None:
        rows = []
        for crop in crops[:50]:
            mask = fallback_mask(*crop.shape[:2])
            if np.sum(mask) > 500:
                rows.append(feature_row(extract_jersey_colors(crop, mask)))
        if len(rows) < 2:
            return
        rows = np.array(rows)
        self.kmeans = KMeans(n_clusters=2, random_state=42)
        labels = self.kmeans.fit_predict(rows)
        ratios = [np.mean(rows[labels == c, 0]) if np.any(labels == c) else 0 for c in range(2)]
        if ratios[1] > ratios[0]:
            self.kmeans.cluster_centers_ = self.kmeans.cluster_centers_[[1, 0]]
        # the reference stores the ratios BEFORE the swap under the swapped names (:219-228); kept as is
        self.team_colors = {0: {"is_white": ratios[0], "name": "Away (White)"}, 1: {"is_white": ratios[1], "name": "Home (Colored)"}}

    def predict(self, crops: List[np.ndarray], tracker_ids: Optional[np.ndarray] = None) -> np.ndarray:
        if not crops:
            return np.array([])
        out = []
        for i, crop in enumerate(crops):
            if self.kmeans is not None:
                f = extract_jersey_colors(crop, fallback_mask(*crop.shape[:2]))
                team = self.kmeans.predict(np.array([feature_row(f)]))[0]
            else:
                team, _ = classify_single_jersey(crop)
            if tracker_ids is not None and i < len(tracker_ids) and tracker_ids[i] is not None:
                tid = int(tracker_ids[i])
                self.player_history[tid].append(team)
                if len(self.player_history[tid]) > self.history_window:
                    self.player_history[tid] = self.player_history[tid][-self.history_window:]
                if len(self.player_history[tid]) >= 3:
                    team = np.argmax(np.bincount(self.player_history[tid]))
            out.append(team)
        return np.array(out)
