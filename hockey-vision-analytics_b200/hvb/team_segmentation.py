"""SegmentationTeamClassifier without GrabCut — the colour features of the reference's default team classifier
(hockey/common/team_segmentation.py:98-292) on the K3a colour pass.

``segment_player`` here always returns the rectangle the reference falls back to when GrabCut raises
(team_segmentation.py:87-96: rows [int(0.2h), int(0.6h)), cols [int(0.3w), int(0.7w))); GrabCut itself is a
sequential CPU algorithm outside the hot path.  Everything downstream of the mask is the reference's
arithmetic: ``extract_jersey_colors`` (LAB white test with its uint8 wrap-around, hue histogram of the non-white
pixels, mean S / V) comes from one ``hvb_jersey_color_stats`` launch over all crops — exact integer counts and
sums, one float64 division per value on the host — and ``fit`` / ``predict`` keep the reference's k-means on the
four features, cluster ordering, and 10-deep per-track majority vote.

Not routed to by ``TeamClassifier``'s default flags (those mirror a reference install where only the hybrid
classifier imports); construct it directly.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _ffi
from .runtime import get_context
from .synth import pack_crops

DEFAULTS = (0.5, 0.0, 0.0, 128.0)          # team_segmentation.py:105-111, fewer than 100 masked pixels


def features_from_raw(raw: np.ndarray) -> np.ndarray:
    """hvb_jersey_raw[n] -> float64[n,4] = (is_white, dominant_hue, saturation, brightness), value for value what
    extract_jersey_colors computes (integer sums are exact, so each entry is a single float64 division)."""
    n = raw["n"].astype(np.int64)
    white = raw["white"].astype(np.int64)
    colored = n - white
    out = np.empty((len(raw), 4), np.float64)
    ok = n >= 100
    many = ok & (colored > 50)
    few = ok & ~many
    safe_n = np.maximum(n, 1).astype(np.float64)
    out[:, 0] = white / safe_n
    out[:, 1] = np.where(many, np.argmax(raw["hue_hist"], axis=1) * 10, 0)
    out[:, 2] = np.where(many, raw["sat_colored"] / np.maximum(colored, 1).astype(np.float64), raw["sat_all"] / safe_n)
    out[:, 3] = raw["val_all"] / safe_n
    out[few, 1] = 0
    out[~ok] = DEFAULTS
    return out


class SegmentationTeamClassifier:
    def __init__(self, device: str = "cuda:0", visualize_segmentation: bool = False):
        self.device = device
        self.visualize_segmentation = visualize_segmentation
        self.ctx = get_context(device)
        self.player_history: Dict[int, List[int]] = defaultdict(list)
        self.history_window = 10
        self.kmeans = None
        self.team_colors = None
        self.last_masks: Dict[int, np.ndarray] = {}

    # ------------------------------------------------------------------ mask + colour features
    def segment_player(self, crop: np.ndarray) -> np.ndarray:
        height, width = crop.shape[:2]
        mask = np.zeros((height, width), dtype=bool)
        mask[int(height * 0.2):int(height * 0.6), int(width * 0.3):int(width * 0.7)] = True
        return mask

    def _raw(self, crops: List[np.ndarray], roi_mode: int = _ffi.ROI_SEGMENT) -> np.ndarray:
        buf, desc = pack_crops(crops)
        cd = np.zeros((len(crops),), _ffi.CROP_DESC)
        cd["offset"], cd["pitch"], cd["h"], cd["w"] = desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3]
        return self.ctx.jersey_color_stats_host(buf, cd, roi_mode)

    def jersey_features(self, crops: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        """(float64[n,4] features, int[n] masked-pixel counts) of all crops in one launch."""
        raw = self._raw(list(crops))
        return features_from_raw(raw), raw["n"].astype(np.int64)

    def jersey_features_from_frames(self, frames_dev, xyxy, frame_idx=None) -> np.ndarray:
        """Device-resident variant: boxes into frames already in HBM (uint8[n,H,W,3]); sv.crop_image geometry."""
        import torch
        n = len(xyxy)
        if n == 0:
            return np.zeros((0, 4), np.float64)
        _, H, W, _ = frames_dev.shape
        boxes = torch.as_tensor(np.asarray(xyxy, np.float32)).to(frames_dev.device)
        fidx = None if frame_idx is None else torch.as_tensor(np.asarray(frame_idx, np.int32)).to(frames_dev.device)
        cd = self.ctx.crops_from_boxes(boxes, fidx, H, W)
        raw = self.ctx.jersey_color_stats(frames_dev, cd, n, _ffi.ROI_SEGMENT)
        return features_from_raw(raw.cpu().numpy()[:n * _ffi.JERSEY_RAW.itemsize].view(_ffi.JERSEY_RAW))

    def extract_jersey_colors(self, crop: np.ndarray, mask: np.ndarray) -> Dict[str, float]:
        """team_segmentation.py:98-148 for a RECTANGULAR mask (what segment_player returns here)."""
        mask = np.asarray(mask, bool)
        rows, cols = np.nonzero(mask.any(1))[0], np.nonzero(mask.any(0))[0]
        if len(rows) == 0:
            f = DEFAULTS
        else:
            t, b, l, r = rows[0], rows[-1] + 1, cols[0], cols[-1] + 1
            if int(mask.sum()) != (b - t) * (r - l):
                raise NotImplementedError("only rectangular masks are supported (GrabCut masks are out of scope)")
            f = features_from_raw(self._raw([crop[t:b, l:r]], _ffi.ROI_WHOLE))[0]
        return {"is_white": f[0], "dominant_hue": int(f[1]), "saturation": f[2], "brightness": f[3]}

    @staticmethod
    def _rule(f) -> Tuple[int, float]:
        if f[0] > 0.4:
            return 0, f[0]
        return 1, min(f[2] / 150, 1.0)

    def classify_single_jersey(self, crop: np.ndarray) -> Tuple[int, float]:
        return self._rule(self.jersey_features([crop])[0][0])

    # ------------------------------------------------------------------ fit / predict (team_segmentation.py:167-292)
    def fit(self, crops: List[np.ndarray], positions: Optional[List[tuple]] = None, frame: Optional[np.ndarray] = None,
            detections=None) -> None:
        from sklearn.cluster import KMeans
        if len(crops) < 10:
            print("Warning: Very few crops for fitting. Results may be unreliable.")
        print(f"Segmenting and analyzing {len(crops)} player crops...")
        feats, npx = self.jersey_features(list(crops[:50]))
        all_features = feats[npx > 500]
        if len(all_features) < 2:
            print("Not enough valid segmentations. Falling back to simple white detection.")
            return
        self.kmeans = KMeans(n_clusters=2, random_state=42)
        labels = self.kmeans.fit_predict(all_features)
        ratios = [np.mean(all_features[labels == c, 0]) if np.any(labels == c) else 0 for c in range(2)]
        if ratios[1] > ratios[0]:
            self.kmeans.cluster_centers_ = self.kmeans.cluster_centers_[[1, 0]]
        print(f"Team 0 (Away/White): avg white ratio = {ratios[0]:.2f}")
        print(f"Team 1 (Home/Colored): avg white ratio = {ratios[1]:.2f}")
        self.team_colors = {0: {"is_white": ratios[0], "name": "Away (White)"},
                            1: {"is_white": ratios[1], "name": "Home (Colored)"}}

    def _vote(self, teams: np.ndarray, tracker_ids) -> np.ndarray:
        out = []
        for i, team in enumerate(teams):
            if tracker_ids is not None and i < len(tracker_ids) and tracker_ids[i] is not None:
                tid = int(tracker_ids[i])
                hist = self.player_history[tid]
                hist.append(int(team))
                if len(hist) > self.history_window:
                    hist = self.player_history[tid] = hist[-self.history_window:]
                if len(hist) >= 3:
                    team = np.argmax(np.bincount(hist))
            out.append(team)
        return np.array(out)

    def predict(self, crops: List[np.ndarray], tracker_ids: Optional[np.ndarray] = None,
                positions: Optional[List[tuple]] = None) -> np.ndarray:
        if not len(crops):
            return np.array([])
        if not self.visualize_segmentation:
            self.last_masks.clear()
        elif self.kmeans is not None and tracker_ids is not None:
            for crop, tid in zip(crops, tracker_ids):
                if tid is not None:
                    self.last_masks[int(tid)] = self.segment_player(crop)
        feats, _ = self.jersey_features(list(crops))
        if self.kmeans is not None:
            teams = self.kmeans.predict(feats)
        else:
            teams = np.array([self._rule(f)[0] for f in feats])
        return self._vote(teams, tracker_ids)

    def predict_from_frames(self, frames_dev, xyxy, frame_idx=None, tracker_ids=None) -> np.ndarray:
        """predict() on boxes into device-resident frames (no host crops)."""
        if not len(xyxy):
            return np.array([])
        feats = self.jersey_features_from_frames(frames_dev, xyxy, frame_idx)
        teams = self.kmeans.predict(feats) if self.kmeans is not None else np.array([self._rule(f)[0] for f in feats])
        return self._vote(teams, tracker_ids)

    def get_segmentation_masks(self, tracker_ids: List[int]) -> Dict[int, np.ndarray]:
        return {tid: self.last_masks[tid] for tid in tracker_ids if tid in self.last_masks}
