"""CPU restatement of the reference's team-classification arithmetic, written against the
same real libraries the reference calls (cv2, Pillow via torchvision.transforms, sklearn).
TEST INFRASTRUCTURE — see oracle/__init__.py.

Follows hockey/common/team_hybrid.py (jersey ROI :49-64, deep features :66-87, colour
features :89-142, fit :155-196, cluster analysis :198-239, predict rule :241-280, temporal vote
:308-328) and hockey/common/team.py (simple HSV rule :76-132).

Pinned against the reference itself: tests/golden/make_golden.py imports the real
team_hybrid.py / team.py from /root/reference in the build container and stores their outputs on
seeded synthetic crops in tests/golden/team_*.npz; tests/test_oracle_team.py checks this module
against those files (and against the live reference import when /root/reference is present).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import cv2
import numpy as np
import torch

N_DEEP = 576
N_COLOR = 49
MNV3_MEAN = (0.485, 0.456, 0.406)
MNV3_STD = (0.229, 0.224, 0.225)


def jersey_rect(h: int, w: int):
    """(top, bottom, left, right) of the jersey ROI inside an h x w crop — team_hybrid.py:49-64.
    Whole crop when h < 40 or w < 20; otherwise Python int() truncation of float products."""
    if h < 40 or w < 20:
        return 0, h, 0, w
    return int(h * 0.1), int(h * 0.6), int(w * 0.2), int(w * 0.8)


def jersey_region(crop: np.ndarray) -> np.ndarray:
    t, b, l, r = jersey_rect(*crop.shape[:2])
    return crop[t:b, l:r]


def simple_rect(h: int, w: int):
    """ROI of TeamClassifier.extract_jersey_region — team.py:76-99 (whole crop if h<30 or w<20,
    or if the sliced region is empty)."""
    if h < 30 or w < 20:
        return 0, h, 0, w
    t, b, l, r = int(h * 0.25), int(h * 0.75), int(w * 0.3), int(w * 0.7)
    if (b - t) * (r - l) == 0:
        return 0, h, 0, w
    return t, b, l, r


def color_stats_raw(roi: np.ndarray) -> Dict[str, np.ndarray]:
    """Exact integer quantities behind the 49-d colour feature of one ROI (real cv2)."""
    hsv = cv2.cvtColor(np.ascontiguousarray(roi), cv2.COLOR_BGR2HSV)
    lab = cv2.cvtColor(np.ascontiguousarray(roi), cv2.COLOR_BGR2LAB)
    hh = cv2.calcHist([hsv], [0], None, [18], [0, 180]).flatten()
    sh = cv2.calcHist([hsv], [1], None, [8], [0, 256]).flatten()
    vh = cv2.calcHist([hsv], [2], None, [8], [0, 256]).flatten()
    hist = np.concatenate([hh, sh, vh]).astype(np.uint32)
    six = np.concatenate([hsv.reshape(-1, 3), lab.reshape(-1, 3)], 1).astype(np.uint64)
    S, V = hsv[..., 1], hsv[..., 2]
    counts = np.array([(S < 30).sum(), (S > 100).sum(), ((V > 200) & (S < 30)).sum()], np.uint32)
    return dict(hist=hist, sums=six.sum(0), sumsq=(six * six).sum(0), counts=counts,
                n=np.uint32(S.size))


def color_features(crops: Sequence[np.ndarray]) -> np.ndarray:
    """float64[n,49] — team_hybrid.py:89-142 (hist float32 / (sum + 1e-7); mean/std /255; ratios)."""
    out = []
    for crop in crops:
        roi = jersey_region(crop)
        hsv = cv2.cvtColor(roi, cv2.COLOR_BGR2HSV)
        lab = cv2.cvtColor(roi, cv2.COLOR_BGR2LAB)
        hists = []
        for ch, bins, hi in ((0, 18, 180), (1, 8, 256), (2, 8, 256)):
            hst = cv2.calcHist([hsv], [ch], None, [bins], [0, hi]).flatten()
            hists.append(hst / (hst.sum() + 1e-7))
        S, V = hsv[:, :, 1], hsv[:, :, 2]
        npx = S.size
        feat = np.concatenate(hists + [
            hsv.mean(axis=(0, 1)) / 255, hsv.std(axis=(0, 1)) / 255,
            lab.mean(axis=(0, 1)) / 255, lab.std(axis=(0, 1)) / 255,
            [np.sum(S < 30) / npx], [np.sum(S > 100) / npx],
            [np.sum((V > 200) & (S < 30)) / npx]])
        out.append(feat)
    return np.array(out)


def make_preprocess():
    """The transforms.Compose of team_hybrid.py:31-36 (real torchvision + Pillow)."""
    from torchvision import transforms
    return transforms.Compose([
        transforms.ToPILImage(),
        transforms.Resize((128, 64)),
        transforms.ToTensor(),
        transforms.Normalize(mean=list(MNV3_MEAN), std=list(MNV3_STD)),
    ])


def preprocess_rois(crops: Sequence[np.ndarray]) -> np.ndarray:
    """float32[n,3,128,64]: the reference's per-crop preprocessing, stacked."""
    pp = make_preprocess()
    return np.stack([pp(jersey_region(c)).numpy() for c in crops]) if len(crops) else \
        np.zeros((0, 3, 128, 64), np.float32)


def deep_features(trunk: torch.nn.Module, crops: Sequence[np.ndarray], device: str = "cpu") -> np.ndarray:
    """float32[n,576] one crop per forward like team_hybrid.py:66-87; failures -> zeros(576)."""
    pp = make_preprocess()
    feats = []
    with torch.no_grad():
        for crop in crops:
            try:
                t = pp(jersey_region(crop)).unsqueeze(0).to(device)
                feats.append(trunk(t).squeeze().cpu().numpy())
            except Exception:
                feats.append(np.zeros(N_DEEP))
    return np.array(feats)


def all_features(trunk, crops, device="cpu") -> np.ndarray:
    """float64[n,625] = hstack(deep float32, colour float64) — team_hybrid.py:144-153."""
    return np.hstack([deep_features(trunk, crops, device), color_features(crops)])


def standardize_fit(features: np.ndarray):
    """StandardScaler.fit_transform semantics (mean, population std, zero variance -> scale 1)."""
    from sklearn.preprocessing import StandardScaler
    sc = StandardScaler()
    return sc, sc.fit_transform(features)


def append_positions(features_normalized: np.ndarray, positions) -> np.ndarray:
    """team_hybrid.py:169-180 — min-max normalised positions x 0.1 appended (D 625 -> 627)."""
    p = np.array(positions)
    pmin, pmax = p.min(axis=0), p.max(axis=0)
    return np.hstack([features_normalized, (p - pmin) / (pmax - pmin + 1e-7) * 0.1])


def rbf_affinity(x: np.ndarray, gamma: float = 1.0):
    """(d2, A) as SpectralClustering(affinity='rbf') builds them (sklearn pairwise_kernels):
    d2 = max(||x||^2 + ||y||^2 - 2 x.y, 0) with an exactly-zero diagonal, A = exp(-gamma d2)."""
    from sklearn.metrics.pairwise import euclidean_distances, rbf_kernel
    x = np.asarray(x, np.float64)
    return euclidean_distances(x, squared=True), rbf_kernel(x, gamma=gamma)


def similarity_rule(features_scaled: np.ndarray) -> np.ndarray:
    """team_hybrid.py:264-280 on standardized features: 0 (white) iff feat[-1] > 0.3 or
    argmax(feat[-10:-7]) == 0, else 1."""
    out = [0 if (f[-1] > 0.3 or np.argmax(f[-10:-7]) == 0) else 1 for f in features_scaled]
    return np.array(out)


def temporal_vote(history: Dict[int, List[int]], predictions: np.ndarray, tracker_ids,
                  window: int = 15, min_hist: int = 5) -> np.ndarray:
    """team_hybrid.py:308-328 (window 15 / >=5) and team.py:282-298 (window 10 / >=3)."""
    out = predictions.copy()
    for i, (p, tid) in enumerate(zip(predictions, tracker_ids)):
        if tid is None:
            continue
        tid = int(tid)
        history[tid].append(p)
        if len(history[tid]) > window:
            history[tid] = history[tid][-window:]
        if len(history[tid]) >= min_hist:
            out[i] = np.argmax(np.bincount(history[tid]))
    return out


def simple_jersey_rule(crop: np.ndarray):
    """TeamClassifier.classify_jersey — team.py:101-132.  Returns (team, confidence)."""
    t, b, l, r = simple_rect(*crop.shape[:2])
    hsv = cv2.cvtColor(crop[t:b, l:r], cv2.COLOR_BGR2HSV)
    bright = np.mean(hsv[:, :, 2])
    sat = np.mean(hsv[:, :, 1])
    white = (hsv[:, :, 2] > 200) & (hsv[:, :, 1] < 30)
    ratio = np.sum(white) / white.size
    if ratio > 0.3 or (bright > 180 and sat < 50):
        return 0, min(ratio * 2, 1.0)
    return 1, min(sat / 150, 1.0)


class HybridReference:
    """Drop-in CPU equivalent of HybridTeamClassifier built from the functions above (used as the
    `--impl reference` arm of bench.py and as the end-to-end oracle)."""

    def __init__(self, trunk: torch.nn.Module, device: str = "cpu", n_clusters: int = 2):
        self.trunk, self.device, self.n_clusters = trunk, device, n_clusters
        self.scaler = None
        self.clusterer = None
        self.cluster_labels = None
        self.affinity_matrix_ = None
        self.player_history: Dict[int, List[int]] = defaultdict(list)
        self.history_window = 15

    def extract_all_features(self, crops):
        return all_features(self.trunk, crops, self.device)

    def fit(self, crops, positions=None, run_clustering: bool = True):
        if len(crops) < self.n_clusters * 2:
            raise ValueError(f"Need at least {self.n_clusters * 2} crops for clustering")
        feats = self.extract_all_features(crops)
        self.scaler, x = standardize_fit(feats)
        if positions and len(positions) == len(crops):
            x = append_positions(x, positions)
        self.features_normalized_ = x
        if run_clustering:
            import warnings
            from sklearn.cluster import SpectralClustering
            self.clusterer = SpectralClustering(n_clusters=self.n_clusters, affinity="rbf", gamma=1.0,
                                                n_init=10, random_state=42)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.cluster_labels = self.clusterer.fit_predict(x)
            self.affinity_matrix_ = self.clusterer.affinity_matrix_
            self._analyze(crops)
        else:
            self.clusterer = "affinity-only"
            self.affinity_matrix_ = rbf_affinity(x, 1.0)[1]

    def _analyze(self, crops):
        sat = {}
        for cid in range(self.n_clusters):
            members = [c for c, l in zip(crops, self.cluster_labels) if l == cid][:20]
            if members:
                sat[cid] = np.mean([np.mean(cv2.cvtColor(jersey_region(c), cv2.COLOR_BGR2HSV)[:, :, 1])
                                    for c in members])
        if len(sat) == 2 and min(sat, key=sat.get) == 1:
            self.cluster_labels = 1 - self.cluster_labels

    def predict(self, crops, tracker_ids=None):
        if not crops:
            return np.array([])
        x = self.scaler.transform(self.extract_all_features(crops))
        pred = similarity_rule(x)
        if tracker_ids is not None:
            pred = temporal_vote(self.player_history, pred, tracker_ids, self.history_window, 5)
        return pred
