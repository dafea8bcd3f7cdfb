"""B200InferenceSlicer — drop-in for ``sv.InferenceSlicer`` on the puck path the reference documents
(README.md:25, CLAUDE.md:55: 640-px slices, 0.2 overlap).

Two ways to use it, same constructor surface as supervision's:

  * compat:  ``B200InferenceSlicer(callback=fn, ...)(frame)`` — `fn(tile) -> Detections` is called per
    tile on host views exactly like supervision does (thread_workers=1 order); move / merge / NMS
    run through the K2b kernel.
  * fast:    ``B200InferenceSlicer(detector=Detector(...), ...)(frame)`` or ``.run_batch(frames)`` —
    tiles never exist on the host: K1b slices + letterboxes every tile of every frame in ONE launch
    into per-shape-class NCHW batches, the YOLO forward runs per class batch, K2a decodes + NMSes
    every tile, hvb_gather_tiles moves/merges per frame and K2b applies the cross-slice NMS.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import _ffi
from .detections import Detections, crop_image
from .detect import Detector


class OverlapFilter:
    NONE = "none"
    NON_MAX_SUPPRESSION = "non_max_suppression"
    NON_MAX_MERGE = "non_max_merge"


def generate_offsets(resolution_wh, slice_wh, overlap_ratio_wh, overlap_wh) -> np.ndarray:
    """Tile rectangles int[n,4] — InferenceSlicer._generate_offset (clipped edge tiles, y-major)."""
    sw, sh = slice_wh
    W, H = resolution_wh
    if overlap_wh is None:
        ow, oh = int(overlap_ratio_wh[0] * sw), int(overlap_ratio_wh[1] * sh)
    else:
        ow, oh = overlap_wh
    xs = np.arange(0, W, sw - ow)
    ys = np.arange(0, H, sh - oh)
    xmin, ymin = np.meshgrid(xs, ys)
    xmax = np.clip(xmin + sw, 0, W)
    ymax = np.clip(ymin + sh, 0, H)
    return np.stack([xmin, ymin, xmax, ymax], -1).reshape(-1, 4)


class B200InferenceSlicer:
    def __init__(self, callback: Optional[Callable[[np.ndarray], Detections]] = None,
                 slice_wh: Tuple[int, int] = (320, 320), overlap_ratio_wh: Optional[Tuple[float, float]] = (0.2, 0.2),
                 overlap_wh: Optional[Tuple[int, int]] = None, overlap_filter: str = OverlapFilter.NON_MAX_SUPPRESSION,
                 iou_threshold: float = 0.5, thread_workers: int = 1, *, detector: Optional[Detector] = None,
                 tile_imgsz: int = 640, uniform_tiles: bool = False, class_agnostic: bool = False,
                 concurrent_classes: Optional[bool] = None):
        if callback is None and detector is None:
            raise ValueError("either `callback` (compat path) or `detector` (device path) is required")
        if overlap_filter == OverlapFilter.NON_MAX_MERGE:
            raise NotImplementedError("NON_MAX_MERGE is not on the reference's path (it documents NMS)")
        self.callback, self.detector = callback, detector
        self.slice_wh, self.overlap_ratio_wh, self.overlap_wh = tuple(slice_wh), overlap_ratio_wh, overlap_wh
        self.overlap_filter, self.iou_threshold = overlap_filter, iou_threshold
        self.thread_workers = thread_workers          # accepted for signature parity; tiles run in slicer order
        self.tile_imgsz, self.uniform_tiles, self.class_agnostic = tile_imgsz, uniform_tiles, class_agnostic
        import os
        if concurrent_classes is None:
            concurrent_classes = os.environ.get("HVB_SLICER_CONCURRENT", "0") == "1"
        self.concurrent_classes = concurrent_classes
        self._side_streams = None
        self.stage_events = None              # set to a list to collect (stage name, CUDA event) marks of the next run_device call

    def _overlap(self) -> Tuple[int, int]:
        if self.overlap_wh is not None:
            return int(self.overlap_wh[0]), int(self.overlap_wh[1])
        return int(self.overlap_ratio_wh[0] * self.slice_wh[0]), int(self.overlap_ratio_wh[1] * self.slice_wh[1])

    # ------------------------------------------------------------------ compat path
    def _call_with_callback(self, image: np.ndarray) -> Detections:
        h, w = image.shape[:2]
        offsets = generate_offsets((w, h), self.slice_wh, self.overlap_ratio_wh, self.overlap_wh)
        parts = []
        for off in offsets:
            det = self.callback(crop_image(image, off))
            if len(det):
                shift = np.array([off[0], off[1], off[0], off[1]])
                det = Detections(xyxy=det.xyxy + shift, confidence=det.confidence, class_id=det.class_id,
                                 tracker_id=det.tracker_id, data=dict(det.data))
            parts.append(det)
        merged = Detections.merge(parts)
        if self.overlap_filter == OverlapFilter.NONE or len(merged) == 0:
            return merged
        ctx = self.detector.ctx if self.detector is not None else None
        if ctx is None:
            from .runtime import get_context
            ctx = get_context()
        keep = ctx.merge_nms_host(merged.xyxy, merged.confidence, None if self.class_agnostic else merged.class_id,
                                  self.iou_threshold, self.class_agnostic)
        return merged[keep]

    # ------------------------------------------------------------------ device path
    def run_device(self, frames_dev: torch.Tensor, sync: bool = True, hook: Optional[str] = "begin"):
        """frames_dev uint8[n,H,W,3] -> (xyxy f64[total,4], conf f32, cls i32, keep u8, seg i32[n+1]) on device.

        sync=True: one small D2H of the per-tile counts sizes the outputs exactly and triggers the large-capacity
        retry for tiles with more than 1024 candidates.  sync=False: nothing is read back — outputs have the full
        capacity n_tiles*max_det rows, only rows [0, seg[-1]) are meaningful, and the per-tile counts are returned
        as a sixth element so the caller can check for overflow (-1) when it reads the results."""
        det = self.detector
        ctx = det.ctx
        n, h, w, _ = frames_dev.shape
        if det.head_hook is not None and hook is not None:      # hook: "begin" next chunk | "repeat" same chunk | None: caller did it
            det.head_hook.begin_chunk(n, repeat=hook == "repeat")
        mode = _ffi.LB_SLICE_UNIFORM if self.uniform_tiles else _ffi.LB_SLICE_EXACT
        plan = det.plan(n, h, w, mode, self.tile_imgsz, self.slice_wh, self._overlap())
        ev = self.stage_events                                   # bench instrumentation: None, or a list to append to

        def mark(name):
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append((name, e))
        mark("start")
        views = plan.class_views(plan.run(frames_dev))
        mark("K1b slice letterbox")
        n_slots = n * plan.tiles_per_frame
        out = (ctx.empty((n_slots, det.max_det, 4), torch.float32), ctx.empty((n_slots, det.max_det), torch.float32),
               ctx.empty((n_slots, det.max_det), torch.int32), torch.zeros((n_slots,), dtype=torch.int32, device=ctx.device))
        states = []
        if self.concurrent_classes and len(views) > 1:
            # The small shape classes (a handful of clipped edge tiles) underfill the GPU: their forwards run on side
            # streams next to the big 640x640 class; the K2a launches stay on the main stream (one shared work area).
            dev = frames_dev.device
            main = torch.cuda.current_stream(dev)
            if self._side_streams is None:
                self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
            order = sorted(range(len(views)), key=lambda c: -views[c].numel())
            heads_of = {}
            fork = torch.cuda.Event()
            fork.record(main)
            for rank, c in enumerate(order):
                if rank == 0:
                    heads_of[c] = det.forward_heads(views[c], c)              # biggest class on the main stream
                    continue
                side = self._side_streams[(rank - 1) % len(self._side_streams)]
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    heads_of[c] = det.forward_heads(views[c], c)
                for hd in (heads_of[c].tensors() if hasattr(heads_of[c], "tensors") else heads_of[c]):
                    hd.record_stream(main)
            for side in self._side_streams:
                main.wait_stream(side)
            for c in range(len(views)):
                meta_h, meta_d = det._meta_dev(plan, c)
                *_, state = det._decode(heads_of[c], meta_h, meta_d, n_slots, out=out)
                states.append(state)
        else:
            for c, x in enumerate(views):
                heads = det.forward_heads(x, c)
                mark("forward")
                meta_h, meta_d = det._meta_dev(plan, c)
                *_, state = det._decode(heads, meta_h, meta_d, n_slots, out=out)
                mark("K2a decode + NMS")
                states.append(state)
        xyxy, cf, cl, cnt = out
        key = ("slot_off", id(plan))
        if key not in det._meta:
            det._meta[key] = ctx.to_device(plan.slot_offsets())
        if sync:
            # overflow (> 1024 candidates in a tile) is rare: one small D2H of the counts decides
            cnt_h = cnt.cpu().numpy()
            if (cnt_h < 0).any():
                for state in states:
                    cnt_h = det._retry_overflow(xyxy, cf, cl, cnt, state, cnt_h)
            total = int(np.maximum(cnt_h, 0).sum())
        else:
            total = n_slots * det.max_det
        g_xyxy, g_conf, g_cls, g_slot, seg = ctx.gather_tiles(xyxy, cf, cl, cnt, det._meta[key], plan.tiles_per_frame, det.max_det)
        if self.overlap_filter == OverlapFilter.NONE or total == 0:
            keep = torch.ones((total,), dtype=torch.uint8, device=ctx.device)
        else:
            keep = ctx.merge_nms(g_xyxy, g_conf, None if self.class_agnostic else g_cls, seg, n, total, self.iou_threshold,
                                 self.class_agnostic)
        mark("gather + K2b cross-slice NMS")
        if sync:
            return g_xyxy[:total], g_conf[:total], g_cls[:total], keep, seg
        return g_xyxy, g_conf, g_cls, keep, seg, cnt

    def run_batch(self, frames) -> List[Detections]:
        det = self.detector
        frames_dev = det.upload(frames)
        xyxy, conf, cls, keep, seg, cnt = self.run_device(frames_dev, sync=False)      # no host round trip mid-pipeline
        seg_h, cnt_h = seg.cpu().numpy(), cnt.cpu().numpy()
        if (cnt_h < 0).any():                       # a tile overflowed the 1024-candidate tier: redo with the retry path
            xyxy, conf, cls, keep, seg = self.run_device(frames_dev, sync=True, hook="repeat")
            seg_h = seg.cpu().numpy()
        total = int(seg_h[-1])
        xyxy_h, conf_h, cls_h, keep_h = (t[:total].cpu().numpy() for t in (xyxy, conf, cls, keep))
        if (keep_h == 0xFF).any():
            raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "merged detections of one frame exceed the on-chip NMS capacity")
        out = []
        for f in range(len(seg_h) - 1):
            lo, hi = int(seg_h[f]), int(seg_h[f + 1])
            k = keep_h[lo:hi].astype(bool)
            out.append(det._to_detections(xyxy_h[lo:hi][k], conf_h[lo:hi][k], cls_h[lo:hi][k]))
            out[-1].xyxy = xyxy_h[lo:hi][k].copy()     # float64 like supervision's moved boxes
        return out

    def __call__(self, image: np.ndarray) -> Detections:
        if self.callback is not None:
            return self._call_with_callback(image)
        return self.run_batch(image)[0]
