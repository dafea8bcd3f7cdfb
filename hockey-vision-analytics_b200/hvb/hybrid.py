"""HybridTeamClassifier — B200 drop-in for the reference class of the same name
(hockey/common/team_hybrid.py:13-328): same constructor, ``fit(crops, positions=None)``,
``predict(crops, tracker_ids=None)``, ``extract_*`` methods, attributes ``scaler``, ``clusterer``,
``cluster_labels``, ``player_history``, ``history_window``.

Where the arithmetic runs:
  jersey ROI + cvtColor HSV/LAB + calcHist + moments -> 49-d colour features      K3a kernel
  jersey ROI + Pillow-exact resize + /255 + Normalize -> float32[n,3,128,64]      K3b kernel
  MobileNetV3-small trunk (batched, fp32, TF32 off)                               PyTorch (backbone only)
  StandardScaler fit/transform                                                    K4a kernels
  RBF affinity (tcgen05 Gram + float64 refinement)                                K4a kernels
  spectral embedding + k-means on the precomputed affinity                        scikit-learn on host (as in the reference);
                                                                                   opt-in spectral="device": dense eigh on the GPU
  similarity rule + temporal vote (a few integers per crop)                       host

The list-of-numpy-views call surface is kept (crops are packed once and uploaded); the
``*_from_frame`` methods are the fast path where crops never exist on the host.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ffi
from .runtime import Context, get_context, nvtx
from .synth import pack_crops

N_DEEP = 576
N_COLOR = 49


class _Scaler:
    """Holds the fitted StandardScaler state (same attribute names as sklearn's)."""

    def __init__(self):
        self.mean_ = None
        self.scale_ = None
        self.var_ = None
        self.n_samples_seen_ = 0
        self._mean_dev = None
        self._scale_dev = None

    def transform(self, x: np.ndarray) -> np.ndarray:
        return (np.asarray(x, np.float64) - self.mean_) / self.scale_


class _TrunkGraph:
    """Static input / output buffers + the captured forward of the MobileNetV3 trunk at one bucketed batch size.
    begin() -> the static input (K3b writes it), replay() -> the static output, end() marks both free again; an event
    orders reuse across streams (the chunk pipeline runs the team stage on its own stream)."""

    def __init__(self, clf: "HybridTeamClassifier", rows: int):
        self.clf, self.rows = clf, rows
        self.graph = self.x = self.y = self.free = None
        self.failed = False

    def ready(self) -> bool:
        """Captured on the first use of a bucket size: one eager forward first (cuDNN autotuning for this batch size), then
        the capture, both on a side stream.  Begun / ended by hand instead of `with torch.cuda.graph(...)`: that context
        manager synchronises the device and empties the caching allocator on entry, which made every later eager forward
        of the process re-cudaMalloc its activations (measured: the eager frame-at-a-time loop fell from 190 to 96
        frames/s).  thread_local: the chunk pipeline's staging thread allocates and copies while this captures."""
        if self.graph is not None:
            return True
        if self.failed:
            return False
        dev = self.clf.ctx.device
        cur = torch.cuda.current_stream(dev)
        x = torch.zeros((self.rows, 3, 128, 64), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.stream(side):
                self.clf._trunk_forward(x)
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    y = self.clf._trunk_forward(x)
                finally:
                    g.capture_end()                # also ends a capture whose body raised (the error then surfaces here)
        except RuntimeError:
            self.failed = True                 # e.g. a capture already in progress on this thread: stay on the eager forward
            return False
        cur.wait_stream(side)
        x.record_stream(cur)
        self.graph, self.x, self.y = g, x, y
        return True

    def begin(self) -> torch.Tensor:
        if self.free is not None:
            torch.cuda.current_stream(self.x.device).wait_event(self.free)
        return self.x

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.y

    def end(self) -> None:
        self.free = torch.cuda.Event()
        self.free.record(torch.cuda.current_stream(self.x.device))


class HybridTeamClassifier:
    def __init__(self, device: str = "cuda:0", n_clusters: int = 2, trunk: Optional[torch.nn.Module] = None,
                 seed: int = 0, affinity_mode: int = 0, fold_batchnorm: bool = True, spectral: str = "sklearn",
                 affinity_gamma=1.0):
        if spectral not in ("sklearn", "device"):
            raise ValueError("spectral must be 'sklearn' (the reference's solver) or 'device'")
        self.ctx: Context = get_context(device)
        self.device = device
        self.n_clusters = n_clusters
        if trunk is None:
            from .models import build_trunk
            trunk = build_trunk(seed)
        import copy
        self.feature_extractor = copy.deepcopy(trunk).to(self.ctx.device).eval()    # the caller's module stays where it is
        if fold_batchnorm:
            from .models.mnv3 import fold_bn
            fold_bn(self.feature_extractor)
        self.scaler = _Scaler()
        self.player_history: Dict[int, List[int]] = defaultdict(list)
        self.history_window = 15
        self.clusterer = None
        self.cluster_labels = None
        self.affinity_matrix_ = None
        self.affinity_mode = affinity_mode      # 0 = tcgen05 Gram + fp64 refine, 1 = fp64 only
        self.spectral = spectral                # "device": dense eigensolver on the GPU (hvb/spectral.py), opt-in
        self.affinity_gamma = affinity_gamma    # 1.0 = the reference; "scale" = 1 / n_features (does not underflow), opt-in
        import os
        self.graph_trunk = os.environ.get("HVB_TRUNK_GRAPH", "1") != "0"      # replay the trunk forward as a CUDA graph per bucket size
        self._trunk_graphs: Dict[int, _TrunkGraph] = {}

    # ------------------------------------------------------------------ geometry (host, O(1))
    def extract_jersey_region(self, crop: np.ndarray) -> np.ndarray:
        h, w = crop.shape[:2]
        if h < 40 or w < 20:
            return crop
        return crop[int(h * 0.1):int(h * 0.6), int(w * 0.2):int(w * 0.8)]

    # ------------------------------------------------------------------ device helpers
    def _upload_crops(self, crops: Sequence[np.ndarray]):
        buf, desc = pack_crops(crops)
        cd = np.zeros((len(crops),), _ffi.CROP_DESC)
        cd["offset"], cd["pitch"], cd["h"], cd["w"] = desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3]
        pix = torch.from_numpy(buf)
        if torch.cuda.is_available():
            pix = pix.pin_memory()
        return pix.to(self.ctx.device, non_blocking=True), self.ctx.struct_to_device(cd)

    def _trunk_forward(self, x: torch.Tensor) -> torch.Tensor:
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad():
                return self.feature_extractor(x).flatten(1)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev

    def _features_device(self, pixels: torch.Tensor, crops_dev: torch.Tensor, n: int, want_raw: bool = False):
        """-> float64[n,625] on the device (deep 0..575 | colour 576..624), optional raw colour stats."""
        ctx = self.ctx
        feats = ctx.empty((n, N_DEEP + N_COLOR), torch.float64)
        raw = None
        with nvtx("hvb:K3a colour features"):
            res = ctx.color_features(pixels, crops_dev, n, _ffi.ROI_HYBRID, want_raw=want_raw,
                                     out_feat=feats[:, N_DEEP:], feat_stride=N_DEEP + N_COLOR)
        if want_raw:
            _, raw = res
        # The number of crops changes from call to call (tracked players of a chunk / a frame); with cudnn.benchmark on,
        # every new batch size would re-run the autotuner for each of the trunk's ~50 convolutions (seconds — measured:
        # the 32-frame drop-in fell to 42 frames/s).  The trunk therefore only ever sees a few bucketed batch sizes: K3b
        # writes the first n rows of a bucket-sized tensor, the padding rows are zero and their outputs are dropped.
        nb = self._bucket(n)
        tg = self._trunk_graph(nb)
        with nvtx("hvb:K3b crop preprocessing"):
            x, valid = ctx.mnv3_preprocess(pixels, crops_dev, n, _ffi.ROI_HYBRID, rows=nb, out=tg.begin() if tg else None)
        with nvtx("hvb:mobilenetv3 trunk"):
            deep = (tg.replay() if tg else self._trunk_forward(x))[:n]
        deep = deep * (valid == 1).to(deep.dtype).unsqueeze(1)       # failed preprocessing -> zeros(576)
        feats[:, :N_DEEP] = deep.to(torch.float64)
        if tg:
            tg.end()
        return feats, raw, valid

    def _trunk_graph(self, rows: int):
        """The trunk forward at bucket size `rows` as ONE CUDA-graph replay (frame-at-a-time calls are launch-bound: the
        ~150 launches of the eager MobileNetV3 forward took 1.72 ms of host time for 11 crops on the B200 box,
        tools/probe_e2e4k.py; the replay 0.3 ms).  Captured on the first call at a bucket size."""
        if not self.graph_trunk:
            return None
        g = self._trunk_graphs.get(rows)
        if g is None:
            self._trunk_graphs[rows] = g = _TrunkGraph(self, rows)
        return g if g.ready() else None

    @staticmethod
    def _bucket(n: int) -> int:
        """Batch size the trunk runs at for n crops: 16, 64, then multiples of 128."""
        if n <= 16:
            return 16
        if n <= 64:
            return 64
        return (n + 127) // 128 * 128

    def _raise_on_empty(self, feats_color_first: torch.Tensor):
        # cv2.cvtColor raises on an empty ROI in the reference (no try/except around the colour path)
        if bool(torch.isnan(feats_color_first).any()):
            raise ValueError("empty jersey region in extract_color_features")

    # ------------------------------------------------------------------ reference-surface extractors
    def extract_deep_features(self, crops: List[np.ndarray]) -> np.ndarray:
        if not len(crops):
            return np.array([])
        pixels, cd = self._upload_crops(crops)
        n = len(crops)
        x, valid = self.ctx.mnv3_preprocess(pixels, cd, n, _ffi.ROI_HYBRID, rows=self._bucket(n))
        deep = self._trunk_forward(x)[:n] * (valid == 1).float().unsqueeze(1)
        return deep.cpu().numpy()

    def extract_color_features(self, crops: List[np.ndarray]) -> np.ndarray:
        if not len(crops):
            return np.array([])
        pixels, cd = self._upload_crops(crops)
        f = self.ctx.color_features(pixels, cd, len(crops), _ffi.ROI_HYBRID)
        self._raise_on_empty(f[:, 0])
        return f.cpu().numpy()

    def extract_all_features(self, crops: List[np.ndarray]) -> np.ndarray:
        pixels, cd = self._upload_crops(crops)
        f, _, _ = self._features_device(pixels, cd, len(crops))
        self._raise_on_empty(f[:, N_DEEP])
        return f.cpu().numpy()

    # ------------------------------------------------------------------ fit
    def fit(self, crops: List[np.ndarray], positions: Optional[List[Tuple[float, float]]] = None) -> None:
        if len(crops) < self.n_clusters * 2:
            raise ValueError(f"Need at least {self.n_clusters * 2} crops for clustering")
        pixels, cd = self._upload_crops(crops)
        feats, raw, _ = self._features_device(pixels, cd, len(crops), want_raw=True)
        self._raise_on_empty(feats[:, N_DEEP])
        raw_h = raw.cpu().numpy().view(_ffi.COLOR_RAW)[: len(crops)]
        self.fit_features(feats, positions, raw_h)

    def fit_features(self, feats: torch.Tensor, positions=None, raw_stats: Optional[np.ndarray] = None,
                     cluster: bool = True) -> None:
        """Fit from a float64[N,625] device feature matrix (e.g. gathered from several GPUs)."""
        ctx = self.ctx
        n = feats.shape[0]
        mean, scale, xs = ctx.standardize(feats)
        sc = self.scaler
        sc._mean_dev, sc._scale_dev = mean, scale
        sc.mean_, sc.scale_ = mean.cpu().numpy(), scale.cpu().numpy()
        sc.var_ = sc.scale_ ** 2
        sc.n_samples_seen_ = n
        if positions and len(positions) == n:
            p = np.array(positions)
            pmin, pmax = p.min(axis=0), p.max(axis=0)
            pn = (p - pmin) / (pmax - pmin + 1e-7) * 0.1
            xs = torch.cat([xs, ctx.to_device(pn.astype(np.float64))], 1).contiguous()
        self.features_normalized_ = xs
        gamma = 1.0 / xs.shape[1] if self.affinity_gamma == "scale" else float(self.affinity_gamma)
        _, a = ctx.gram_affinity(xs, gamma, self.affinity_mode, want_d2=False, want_a=True)
        if not cluster:                      # scaler + affinity only (multi-GPU ranks other than the one that clusters)
            self.affinity_matrix_dev_ = a
            self.clusterer = "affinity-only"
            return
        self.affinity_matrix_ = a.cpu().numpy()
        import warnings
        if self.spectral == "device":
            from .spectral import DeviceSpectralClustering
            self.clusterer = DeviceSpectralClustering(n_clusters=self.n_clusters, n_init=10, random_state=42)
            self.clusterer.affinity_matrix_ = self.affinity_matrix_
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.cluster_labels = self.clusterer.fit_predict(a)
        else:
            from sklearn.cluster import SpectralClustering
            self.clusterer = SpectralClustering(n_clusters=self.n_clusters, affinity="precomputed", n_init=10, random_state=42)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.cluster_labels = self.clusterer.fit_predict(self.affinity_matrix_)
        if raw_stats is not None:
            self._analyze_clusters(raw_stats, self.cluster_labels)

    def _analyze_clusters(self, raw: np.ndarray, labels: np.ndarray) -> None:
        """team_hybrid.py:198-239: cluster with the lower mean saturation (first 20 crops) becomes 0."""
        stats = {}
        for cid in range(self.n_clusters):
            idx = np.nonzero(labels == cid)[0][:20]
            if len(idx):
                sat = raw["sums"][idx, 1].astype(np.float64) / raw["n"][idx]
                white = raw["counts"][idx, 2].astype(np.float64) / raw["n"][idx]
                stats[cid] = dict(avg_saturation=float(np.mean(sat)), avg_white_ratio=float(np.mean(white)),
                                  count=int((labels == cid).sum()))
        self.cluster_stats_ = stats
        if len(stats) == 2 and min(stats, key=lambda k: stats[k]["avg_saturation"]) == 1:
            self.cluster_labels = 1 - self.cluster_labels

    # ------------------------------------------------------------------ predict
    def predict(self, crops: List[np.ndarray], tracker_ids: Optional[np.ndarray] = None) -> np.ndarray:
        if not len(crops):
            return np.array([])
        pixels, cd = self._upload_crops(crops)
        return self._predict_device(pixels, cd, len(crops), tracker_ids)

    def predict_from_frame(self, frames_dev: torch.Tensor, xyxy: torch.Tensor, frame_idx: Optional[torch.Tensor] = None,
                           tracker_ids: Optional[np.ndarray] = None) -> np.ndarray:
        """Fast path: frames already resident on the GPU, crops described by detection boxes."""
        n = xyxy.shape[0]
        if n == 0:
            return np.array([])
        h, w = frames_dev.shape[-3], frames_dev.shape[-2]
        cd = self.ctx.crops_from_boxes(xyxy.to(self.ctx.device, torch.float32), frame_idx, h, w)
        return self._predict_device(frames_dev, cd, n, tracker_ids)

    def features_from_frame(self, frames_dev: torch.Tensor, xyxy: torch.Tensor, frame_idx: Optional[torch.Tensor] = None):
        n = xyxy.shape[0]
        h, w = frames_dev.shape[-3], frames_dev.shape[-2]
        cd = self.ctx.crops_from_boxes(xyxy.to(self.ctx.device, torch.float32), frame_idx, h, w)
        return self._features_device(frames_dev, cd, n, want_raw=True)

    def _predict_device(self, pixels, cd, n, tracker_ids):
        feats, raw, _ = self._features_device(pixels, cd, n, want_raw=self.clusterer is None)
        if self.clusterer is None:
            self._raise_on_empty(feats[:, N_DEEP])
            r = raw.cpu().numpy().view(_ffi.COLOR_RAW)[:n]
            sat = r["sums"][:, 1].astype(np.float64) / r["n"]
            white = r["counts"][:, 2].astype(np.float64) / r["n"]
            pred = np.where((white > 0.25) | (sat < 40), 0, 1)          # _simple_classify, :282-306
        else:
            xs = self.ctx.scale_transform(feats, self.scaler._mean_dev, self.scaler._scale_dev)
            # one device -> host read: the entries the rule reads + the first colour feature (NaN = empty jersey region)
            tail = torch.cat([xs[:, -10:], feats[:, N_DEEP:N_DEEP + 1]], 1).cpu().numpy()
            if np.isnan(tail[:, 10]).any():
                raise ValueError("empty jersey region in extract_color_features")
            tail = tail[:, :10]
            pred = np.array([0 if (t[-1] > 0.3 or np.argmax(t[0:3]) == 0) else 1 for t in tail])
        if tracker_ids is not None:
            pred = self._apply_temporal_consistency(pred, tracker_ids)
        return pred

    def _apply_temporal_consistency(self, predictions: np.ndarray, tracker_ids) -> np.ndarray:
        out = predictions.copy()
        for i, (p, tid) in enumerate(zip(predictions, tracker_ids)):
            if tid is None:
                continue
            tid = int(tid)
            hist = self.player_history[tid]
            hist.append(p)
            if len(hist) > self.history_window:
                self.player_history[tid] = hist = hist[-self.history_window:]
            if len(hist) >= 5:
                out[i] = np.argmax(np.bincount(hist))
        return out
