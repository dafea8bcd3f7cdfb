"""Chunked hot path: what ``VideoProcessor.process_frame`` does between "decoded BGR frame" and
"detections + team ids" (reference hockey/main.py:259-281), run on a chunk of frames per launch so
that every libhvb kernel sees enough work to be bandwidth- rather than latency-bound.

    frames --H2D--> K1a letterbox --> YOLOv8 forward (torch) --> K2a decode+NMS
                 \\-> boxes --> K3a colour features + K3b crop preprocessing --> MobileNetV3 (torch)
                               --> K4a scale_transform --> similarity rule (+ temporal vote on host)

ByteTrack and the temporal vote are order-dependent per clip and stay on the clip's owner (host).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _ffi
from .detect import Detector, PLAYER_CLASS_ID
from .hybrid import HybridTeamClassifier
from .runtime import get_context
from .slicer import B200InferenceSlicer


def _stream_state(owner, dev):
    """The staging state of a process_stream pipeline, kept on its owner between calls: the copy stream, the two device
    frame buffers (+ the events that guard their reuse) and the three sets of pinned result buffers.  Measured on the
    B200 box (tools/probe_e2e4k.py): creating them per call — cudaHostAlloc of the result buffers, cudaMalloc of 2 x 400 MB
    under a fresh stream — cost ~90 ms, more than three 4K chunks."""
    st = getattr(owner, "_stream_state_", None)
    if st is None or st[0] != str(dev):
        st = (str(dev), torch.cuda.Stream(device=dev), [None, None], [None, None], [dict(), dict(), dict()])
        owner._stream_state_ = st
    return st[1:]


class HotPath:
    def __init__(self, device="cuda:0", yolo_scale: str = "m", nc: int = 2, imgsz: int = 1280, conf: float = 0.4,
                 seed: int = 0, trunk: Optional[torch.nn.Module] = None, affinity_mode: int = 0, fuse: bool = True,
                 channels_last: bool = True, autocast_dtype: Optional[torch.dtype] = None, overlap_team: bool = True):
        from .models import build_trunk, build_yolov8
        self.ctx = get_context(device)
        import os
        if os.environ.get("HVB_CUDNN_BENCHMARK_LIMIT") is not None:      # 0 = let cuDNN's autotuner try every algorithm
            torch.backends.cudnn.benchmark_limit = int(os.environ["HVB_CUDNN_BENCHMARK_LIMIT"])
        # The team stage (K3a/K3b, MobileNetV3, K4a: ~170 small launches, 2 ms) underfills the GPU; when its boxes are
        # already known (tracker output / previous chunk) it runs on a high-priority side stream next to the detection
        # stage (large HBM-bound kernels) instead of after it.
        self.overlap_team = overlap_team
        self._side = None
        torch.backends.cudnn.benchmark = True
        self.detector = Detector(build_yolov8(yolo_scale, nc, seed), device, imgsz=imgsz, conf=conf,
                                 class_names={0: "player", 1: "goalie"}, fuse=fuse, channels_last=channels_last,
                                 autocast_dtype=autocast_dtype)
        self.classifier = HybridTeamClassifier(device=device, trunk=trunk if trunk is not None else build_trunk(seed, calibrate=True),
                                               affinity_mode=affinity_mode)

    def classifier_router(self):
        """A TeamClassifier (the reference's router class) that routes to this path's fitted hybrid classifier."""
        from .team import TeamClassifier
        tc = TeamClassifier.__new__(TeamClassifier)
        TeamClassifier.__init__(tc, device=str(self.ctx.device), use_hybrid=False)
        tc.use_hybrid, tc.hybrid_classifier = True, self.classifier
        return tc

    # ------------------------------------------------------------------ one-off fit (per clip)
    def fit_from_frames(self, frames_dev: torch.Tensor, boxes: torch.Tensor, frame_idx: torch.Tensor):
        feats, raw, _ = self.classifier.features_from_frame(frames_dev, boxes, frame_idx)
        raw_h = raw.cpu().numpy().view(_ffi.COLOR_RAW)[: boxes.shape[0]]
        self.classifier.fit_features(feats, None, raw_h)
        return feats

    # ------------------------------------------------------------------ per-chunk hot path
    def detect_device(self, frames_dev: torch.Tensor):
        return self.detector.detect_device(frames_dev)

    def team_device(self, frames_dev: torch.Tensor, boxes: torch.Tensor, frame_idx: torch.Tensor) -> torch.Tensor:
        """Scaled tail features float64[M,10] (what the similarity rule reads), on the device."""
        clf = self.classifier
        m = boxes.shape[0]
        h, w = frames_dev.shape[1], frames_dev.shape[2]
        cd = self.ctx.crops_from_boxes(boxes, frame_idx, h, w)
        feats, _, _ = clf._features_device(frames_dev, cd, m)
        xs = self.ctx.scale_transform(feats, clf.scaler._mean_dev, clf.scaler._scale_dev)
        return xs[:, -10:].contiguous()

    def process_chunk_device(self, frames_dev: torch.Tensor, team_boxes: Optional[torch.Tensor] = None,
                             team_frame_idx: Optional[torch.Tensor] = None):
        """Everything on the device; returns the tensors a caller would copy back (+ `state`, what an overflow retry
        of K2a needs: the head tensors and the per-image meta)."""
        if team_boxes is not None and self.overlap_team and team_boxes.shape[0]:
            dev = frames_dev.device
            main = torch.cuda.current_stream(dev)
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev, priority=-1)
            side = self._side
            side.wait_stream(main)                                    # frames / boxes are ready at this point of main
            with torch.cuda.stream(side):
                tail = self.team_device(frames_dev, team_boxes, team_frame_idx)
            for t in (frames_dev, team_boxes, team_frame_idx):
                if t is not None:
                    t.record_stream(side)
            xyxy, conf, cls, cnt, state = self.detect_device(frames_dev)
            main.wait_stream(side)                                    # team stage finished long before detection does
            tail.record_stream(main)
            return dict(xyxy=xyxy, conf=conf, cls=cls, count=cnt, team_tail=tail, team_frame_idx=team_frame_idx, state=state,
                        team_from_detections=False)
        xyxy, conf, cls, cnt, state = self.detect_device(frames_dev)
        out = dict(xyxy=xyxy, conf=conf, cls=cls, count=cnt, state=state, team_from_detections=team_boxes is None)
        self._team_of(out, frames_dev, team_boxes, team_frame_idx)
        return out

    def _team_of(self, out, frames_dev, team_boxes, team_frame_idx):
        if team_boxes is None:
            # boxes of detected players (class 0), compacted with torch indexing (tiny)
            conf, cls, cnt, xyxy = out["conf"], out["cls"], out["count"], out["xyxy"]
            n, md = conf.shape
            valid = (torch.arange(md, device=conf.device)[None, :] < cnt[:, None]) & (cls == PLAYER_CLASS_ID)
            fi, ki = torch.nonzero(valid, as_tuple=True)
            team_boxes = xyxy[fi, ki].contiguous()
            team_frame_idx = fi.to(torch.int32)
        out["team_tail"] = self.team_device(frames_dev, team_boxes, team_frame_idx) if team_boxes.shape[0] else \
            torch.zeros((0, 10), dtype=torch.float64, device=frames_dev.device)
        out["team_frame_idx"] = team_frame_idx

    def _resolve_overflow(self, out, frames_dev, cnt_host: np.ndarray) -> np.ndarray:
        """Frames that reported -1 (> 1024 candidates above conf) are re-run with K2a's 8192-candidate tier, like
        Detector.detect_batch does (ultralytics' NMS, max_nms = 30000, never fails here); raises only if that tier
        overflows too.  The team stage is redone when its boxes came from the detections."""
        cnt_host = self.detector._retry_overflow(out["xyxy"], out["conf"], out["cls"], out["count"], out["state"], cnt_host)
        if out["team_from_detections"]:
            self._team_of(out, frames_dev, None, None)
        return cnt_host

    @staticmethod
    def rule(tail: np.ndarray) -> np.ndarray:
        """HybridTeamClassifier._predict_by_similarity on the scaled tail (team_hybrid.py:264-280)."""
        if len(tail) == 0:
            return np.zeros((0,), np.int64)
        return np.where((tail[:, -1] > 0.3) | (np.argmax(tail[:, 0:3], axis=1) == 0), 0, 1).astype(np.int64)

    def process_chunk(self, frames: np.ndarray, team_boxes: Optional[np.ndarray] = None,
                      team_frame_idx: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        """Public host API: host frames in (pinned H2D inside), host results out."""
        det = self.detector
        frames_dev = det.upload(frames)
        tb = ti = None
        if team_boxes is not None:
            tb = torch.from_numpy(np.ascontiguousarray(team_boxes, np.float32)).to(self.ctx.device, non_blocking=True)
            ti = torch.from_numpy(np.ascontiguousarray(team_frame_idx, np.int32)).to(self.ctx.device, non_blocking=True)
        out = self.process_chunk_device(frames_dev, tb, ti)
        cnt = out["count"].cpu().numpy()
        if (cnt < 0).any():
            cnt = self._resolve_overflow(out, frames_dev, cnt)
        res = dict(count=cnt, xyxy=out["xyxy"].cpu().numpy(), conf=out["conf"].cpu().numpy(), cls=out["cls"].cpu().numpy())
        res["team"] = self.rule(out["team_tail"].cpu().numpy())
        return res


    def process_stream(self, chunks):
        """Pipelined host API: `chunks` yields (frames uint8[n,H,W,3] pinned torch tensor or numpy, team_boxes
        float32[M,4], team_frame_idx int32[M]) per step; results are yielded in order, one step behind.

        Three things overlap: the H2D copy of chunk i (side stream, two device frame buffers guarded by
        events), the kernels of chunk i-1 (main stream), and the host reading chunk i-1's results (asynchronous
        D2H into pinned buffers + an event).  The host only ever blocks on the *previous* chunk, so the GPU
        always has the next chunk's work queued behind the current one."""
        dev = self.ctx.device
        main = torch.cuda.current_stream(dev)
        copy_stream, bufs, free_ev, pinned_out = _stream_state(self, dev)

        def stage(i, item):
            frames, tb, ti = item
            src = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames)).pin_memory()
            slot = i & 1
            with torch.cuda.stream(copy_stream):
                if bufs[slot] is None or bufs[slot].shape != src.shape:
                    # allocated UNDER the copy stream: a block handed out under the main stream may be one that main-stream
                    # kernels still in flight are using (the caching allocator reuses host-freed blocks stream-ordered),
                    # and the copy below would overwrite it concurrently
                    bufs[slot] = torch.empty(src.shape, dtype=torch.uint8, device=dev)
                if free_ev[slot] is not None:
                    copy_stream.wait_event(free_ev[slot])            # previous user of this buffer is done
                bufs[slot].copy_(src, non_blocking=True)
                tbd = torch.from_numpy(np.ascontiguousarray(tb, np.float32)).pin_memory().to(dev, non_blocking=True)
                tid = torch.from_numpy(np.ascontiguousarray(ti, np.int32)).pin_memory().to(dev, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            return slot, tbd, tid, ready, src

        def launch(i, staged):
            slot, tbd, tid, ready, _keep = staged
            main.wait_event(ready)
            tbd.record_stream(main); tid.record_stream(main); bufs[slot].record_stream(main)
            out = self.process_chunk_device(bufs[slot], tbd, tid)
            ev = torch.cuda.Event()
            ev.record(main)
            free_ev[slot] = ev
            host = pinned_out[i % 3]
            for k in ("count", "xyxy", "conf", "cls", "team_tail"):
                t = out[k]
                if k not in host or host[k].shape != t.shape or host[k].dtype != t.dtype:
                    host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                host[k].copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            return host, done, _keep, out, bufs[slot]

        def collect(item):
            host, done, _keep, out, frames_dev = item
            done.synchronize()
            cnt = host["count"].numpy().copy()
            if (cnt < 0).any():
                # rare: redo K2a for the overflowing frames with the large tier (the chunk's heads and frame buffer are
                # still alive: the buffer is only recycled two chunks later) and read this chunk's results again
                cnt = self._resolve_overflow(out, frames_dev, cnt)
                for k in ("xyxy", "conf", "cls", "team_tail"):
                    if host[k].shape != out[k].shape:
                        host[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
                    host[k].copy_(out[k])
            res = dict(count=cnt, xyxy=host["xyxy"].numpy().copy(), conf=host["conf"].numpy().copy(), cls=host["cls"].numpy().copy())
            res["team"] = self.rule(host["team_tail"].numpy())
            return res

        inflight = []
        for i, item in enumerate(chunks):
            inflight.append(launch(i, stage(i, item)))
            if len(inflight) > 1:
                yield collect(inflight.pop(0))
        while inflight:
            yield collect(inflight.pop(0))


class SlicedPuckPath:
    """4K puck detection through the slicer: K1b -> YOLOv8n forward per shape class -> K2a -> gather -> K2b."""

    def __init__(self, device="cuda:0", yolo_scale: str = "n", nc: int = 1, conf: float = 0.4, seed: int = 0,
                 uniform_tiles: bool = False, iou_threshold: float = 0.1, fuse: bool = True, channels_last: bool = True,
                 autocast_dtype: Optional[torch.dtype] = None):
        from .models import build_yolov8
        torch.backends.cudnn.benchmark = True
        self.detector = Detector(build_yolov8(yolo_scale, nc, seed), device, imgsz=640, conf=conf, class_names={0: "puck"},
                                 fuse=fuse, channels_last=channels_last, autocast_dtype=autocast_dtype)
        self.slicer = B200InferenceSlicer(detector=self.detector, slice_wh=(640, 640), overlap_ratio_wh=(0.2, 0.2),
                                          iou_threshold=iou_threshold, tile_imgsz=640, uniform_tiles=uniform_tiles)
        self._graphs = {}

    def process_chunk_device(self, frames_dev: torch.Tensor, sync: bool = False, graph: bool = False):
        """Device-resident chunk: no host round trip unless sync=True (exact-size outputs + overflow retry).
        graph=True replays the whole chunk (K1b, every shape-class forward, K2a, gather, K2b: ~900 launches) as ONE
        CUDA graph captured on first use for this chunk shape — the sliced path is launch-bound otherwise."""
        if not graph or sync:
            return self.slicer.run_device(frames_dev, sync=sync)
        key = tuple(frames_dev.shape) + (id(self.detector.head_hook),)
        if self.detector.head_hook is not None:                # outside the graph: stage this chunk's planted tables
            self.detector.head_hook.begin_chunk(frames_dev.shape[0])
        if key not in self._graphs:
            from .runtime import GraphedStep
            self._graphs[key] = GraphedStep(self.detector.ctx, lambda f: self.slicer.run_device(f, sync=False, hook=None), [frames_dev])
        return self._graphs[key](frames_dev)

    def process_chunk(self, frames: np.ndarray):
        return self.slicer.run_batch(frames)

    def process_stream(self, chunks, graph: bool = True):
        """Pipelined host API of the sliced path: `chunks` yields pinned uint8[n,H,W,3] host tensors (or numpy arrays);
        yields per chunk a list of Detections (one per frame), one step behind.  The H2D copy of chunk i (side stream,
        two device buffers) overlaps the kernels of chunk i-1 (one CUDA-graph replay) and the host's read of chunk i-1's
        results (asynchronous D2H into pinned buffers + an event)."""
        det = self.detector
        dev = det.ctx.device
        main = torch.cuda.current_stream(dev)
        copy_stream, bufs, free_ev, pinned_out = _stream_state(self, dev)
        names = ("xyxy", "conf", "cls", "keep", "seg", "count")

        def stage(i, frames):
            src = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames)).pin_memory()
            slot = i & 1
            with torch.cuda.stream(copy_stream):
                if bufs[slot] is None or bufs[slot].shape != src.shape:
                    bufs[slot] = torch.empty(src.shape, dtype=torch.uint8, device=dev)     # under the copy stream (see HotPath.process_stream)
                if free_ev[slot] is not None:
                    copy_stream.wait_event(free_ev[slot])
                bufs[slot].copy_(src, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            return slot, ready, src

        def launch(i, staged):
            slot, ready, keep_alive = staged
            main.wait_event(ready)
            bufs[slot].record_stream(main)
            out = self.process_chunk_device(bufs[slot], graph=graph)
            ev = torch.cuda.Event()
            ev.record(main)
            free_ev[slot] = ev
            host = pinned_out[i % 3]
            for k, t in zip(names, out):
                if k not in host or host[k].shape != t.shape or host[k].dtype != t.dtype:
                    host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                host[k].copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            return host, done, bufs[slot], keep_alive

        def collect(item):
            host, done, frames_dev, _keep = item
            done.synchronize()
            seg, cnt = host["seg"].numpy(), host["count"].numpy()
            if (cnt < 0).any():                                # > 1024 candidates in a tile: redo this chunk with the retry tier
                return self.slicer.run_batch(frames_dev)          # (planted benchmarks never overflow: the hook is not re-armed here)
            total = int(seg[-1])
            xyxy, conf, cls, keep = (host[k].numpy()[:total] for k in ("xyxy", "conf", "cls", "keep"))
            if (keep == 0xFF).any():
                raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "merged detections of one frame exceed the on-chip NMS capacity")
            res = []
            for f in range(len(seg) - 1):
                lo, hi = int(seg[f]), int(seg[f + 1])
                k = keep[lo:hi].astype(bool)
                d = det._to_detections(xyxy[lo:hi][k], conf[lo:hi][k], cls[lo:hi][k])
                d.xyxy = xyxy[lo:hi][k].copy()
                res.append(d)
            return res

        inflight = []
        for i, frames in enumerate(chunks):
            inflight.append(launch(i, stage(i, frames)))
            if len(inflight) > 1:
                yield collect(inflight.pop(0))
        while inflight:
            yield collect(inflight.pop(0))
