"""Detection front end: the drop-in for ``VideoProcessor.detect_players`` (reference
hockey/main.py:177-195) and for the per-tile callback of ``sv.InferenceSlicer``.

    frame(s) --H2D--> K1 letterbox kernel --> torch YOLOv8 forward --> K2a decode+NMS kernel --D2H--> Detections

Only the backbone forward runs in PyTorch; letterboxing, head decode, NMS and box rescaling are
libhvb kernels.  Batches of frames go through one K1 launch and one K2a launch.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ffi
from .detections import Detections
from .runtime import Context, LetterboxPlan, SplitHeads, get_context, nvtx

PLAYER_CLASS_ID = 0
GOALKEEPER_CLASS_ID = 1


def _as_frames(frames) -> np.ndarray | torch.Tensor:
    if isinstance(frames, np.ndarray) and frames.ndim == 3:
        frames = frames[None]
    if isinstance(frames, torch.Tensor) and frames.ndim == 3:
        frames = frames[None]
    return frames


class Detector:
    """YOLOv8 detector whose pre/post-processing runs in libhvb.

    `model` is a module returning the three raw Detect tensors (hvb.models.YOLOv8).  Parameters
    mirror the ultralytics predict call at hockey/main.py:179-184: imgsz, conf (+ iou 0.7,
    max_det 300, class-aware NMS defaults)."""

    def __init__(self, model: torch.nn.Module, device="cuda:0", imgsz: int = 1280, conf: float = 0.4, iou: float = 0.7,
                 max_det: int = 300, agnostic_nms: bool = False, class_names: Optional[Dict[int, str]] = None,
                 autocast_dtype: Optional[torch.dtype] = None, channels_last: bool = False, fuse: bool = False,
                 glue: Optional[bool] = None, cuda_graph: Optional[bool] = False, exact_silu: bool = False):
        self.ctx: Context = get_context(device)
        self.device = self.ctx.device
        if fuse:
            # ultralytics folds BatchNorm into the convolutions before inference (model.fuse()); do the same
            import copy
            from .models.yolov8 import fuse_conv_bn
            model = fuse_conv_bn(copy.deepcopy(model))
        self.model = model.to(self.device).eval()
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.channels_last = channels_last
        self.nc = int(model.nc)
        # K5: run everything between the convolutions (bias, SiLU, residual, concat, upsample, layer 0) in libhvb
        from .models.yolov8 import YOLOv8
        if glue is None:
            glue = fuse and channels_last and autocast_dtype is None and isinstance(model, YOLOv8)
        self.runner = None
        if glue:
            from .models.fused import FusedYOLOv8
            # exact_silu=True: expf + IEEE division (torch's formula) instead of the approximate-unit SiLU (<= 1e-6 relative)
            self.runner = FusedYOLOv8(self.model, self.ctx, exact_silu=exact_silu)
        self.imgsz, self.conf, self.iou, self.max_det, self.agnostic = imgsz, conf, iou, max_det, agnostic_nms
        self.autocast_dtype = autocast_dtype
        self.class_names = class_names or {i: str(i) for i in range(self.nc)}
        # cuda_graph=True: the whole device side of a detect call (K1a, ~270 backbone launches, K2a) is captured once per
        # input shape and replayed — for frame-at-a-time use (process_frame) the eager path is launch-bound.
        # None = automatic: on when the forward is the K5 runner (known to be capturable: no host synchronisation inside).
        self.cuda_graph = (self.runner is not None) if cuda_graph is None else bool(cuda_graph)
        self._graphs: Dict[Tuple, object] = {}
        self._staging: Dict[Tuple, dict] = {}
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        # worker threads of the host-side staging copy: the cores this process may run on, at most 8 (the copy is a short
        # burst per chunk; with 4 cores per rank at N = 8, two threads needed as long as the GPU step: 37 ms per 398 MB)
        self.staging_threads = int(os.environ.get("HVB_STAGING_THREADS", max(1, min(8, cores))))
        self._plans: Dict[Tuple, LetterboxPlan] = {}
        self._meta: Dict[Tuple, torch.Tensor] = {}
        # Synthetic-input hook (hvb.synth.DeviceOverlay): an object with begin_chunk(n_frames) — called once per detect
        # call, outside any CUDA graph — and __call__(heads, cls_index) -> heads, applied to the raw head tensors between
        # the forward and K2a.  None in production.
        self.head_hook = None

    # -------------------------------------------------------------- plumbing
    def plan(self, n: int, h: int, w: int, mode: int = _ffi.LB_WHOLE, imgsz: Optional[int] = None,
             slice_wh=(640, 640), overlap_wh=(128, 128)) -> LetterboxPlan:
        key = (n, h, w, mode, imgsz or self.imgsz, tuple(slice_wh), tuple(overlap_wh))
        if key not in self._plans:
            self._plans[key] = self.ctx.letterbox_plan(n, h, w, mode, imgsz or self.imgsz, True, 32, slice_wh, overlap_wh)
        return self._plans[key]

    def upload(self, frames) -> torch.Tensor:
        """Host frames -> device uint8[n,H,W,3]; device tensors pass through.  `frames` is one frame, an array
        [n,H,W,3] or a sequence of equally sized frames (numpy views are fine).  They are copied frame by frame into a
        PERSISTENT pinned staging buffer (one per shape, two alternating so that the previous asynchronous H2D copy is
        never overwritten) — pinning a fresh 200 MB buffer per chunk costs more than the whole GPU step."""
        if isinstance(frames, torch.Tensor):
            return _as_frames(frames).to(self.device).contiguous()
        if isinstance(frames, np.ndarray):
            frames = _as_frames(frames)
        n = len(frames)
        h, w = frames[0].shape[:2]
        direct = self._upload_pinned(frames, n, h, w)
        if direct is not None:
            return direct
        key = (n, h, w)
        slot = self._staging.pop(key, None)
        if slot is None:
            while len(self._staging) >= 4:                  # keep the pinned buffers of the four most recent shapes only
                old = self._staging.pop(next(iter(self._staging)))
                for e in old["events"]:
                    if e is not None:
                        e.synchronize()
            slot = {"bufs": [torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(2)],
                    "events": [None, None], "next": 0}
        self._staging[key] = slot                           # most recently used last
        i = slot["next"]
        slot["next"] = i ^ 1
        if slot["events"][i] is not None:
            slot["events"][i].synchronize()                 # the H2D copy that last used this buffer has finished
        buf = slot["bufs"][i]
        self._stage(frames, buf, n, h, w)
        dev = buf.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot["events"][i] = ev
        return dev

    def _upload_pinned(self, frames, n: int, h: int, w: int):
        """Frames that already live in page-locked host memory (a decoder writing into buffers from hvb_host_alloc /
        torch's pin_memory, or memory the caller registered) skip the staging copy: one asynchronous H2D copy per frame
        straight from where they lie, then the call waits for those copies — so the caller may reuse its buffers as soon as
        upload returns, exactly as with the staged path.  Returns None when the frames are pageable / strided."""
        rows = [frames[k] for k in range(n)]
        if not all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.shape == (h, w, 3) and f.flags.c_contiguous for f in rows):
            return None
        try:
            if not torch.from_numpy(rows[0]).is_pinned():    # the common case (pageable frames) costs one pointer query
                return None
            if not all(torch.from_numpy(f).is_pinned() for f in rows[1:]):
                return None
        except (RuntimeError, TypeError):                    # read-only arrays etc.: the staged path copes
            return None
        dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        each = h * w * 3
        base = rows[0].ctypes.data
        if all(f.ctypes.data == base + k * each for k, f in enumerate(rows)):     # consecutive frames of one block: one copy
            import ctypes
            whole = np.ctypeslib.as_array((ctypes.c_uint8 * (n * each)).from_address(base)).reshape(n, h, w, 3)
            dev.copy_(torch.from_numpy(whole), non_blocking=True)
        else:
            for k, f in enumerate(rows):
                dev[k].copy_(torch.from_numpy(f), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        return dev

    def _stage(self, frames, buf: torch.Tensor, n: int, h: int, w: int) -> None:
        """Host frames -> the pinned staging buffer through hvb_stage_frames (worker threads inside libhvb, GIL released).
        torch's per-frame copy_ ran single-threaded when called from the staging thread of process_video_chunked: 398 MB
        per 64-frame chunk took as long as the GPU step (tools/probe_staging.py)."""
        import ctypes
        each = h * w * 3
        rows = [frames[k] for k in range(n)]
        if all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.shape == (h, w, 3) and f.flags.c_contiguous for f in rows):
            ptrs = (ctypes.c_void_p * n)(*[f.ctypes.data for f in rows])
            _ffi.check(_ffi.lib().hvb_stage_frames(ctypes.cast(ptrs, ctypes.c_void_p), n, each, buf.data_ptr(), self.staging_threads))
            return
        for k, f in enumerate(rows):                        # strided views / other dtypes: the general copy
            buf[k].copy_(torch.from_numpy(np.ascontiguousarray(f, np.uint8)))

    def forward_heads(self, x: torch.Tensor, cls_index: int = 0) -> List[torch.Tensor]:
        """Backbone forward (PyTorch).  Returns the 3 raw head tensors (SplitHeads from the K5 runner, else float32
        tensors with jointly contiguous spatial dims).  `cls_index`: shape class of the batch (slicer), for head_hook."""
        heads = self._forward_heads(x)
        if self.head_hook is not None:
            heads = self.head_hook(heads, cls_index)
        return heads

    def _forward_heads(self, x: torch.Tensor):
        if self.runner is not None:
            return self.runner(x)
        with torch.no_grad():
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            if self.autocast_dtype is not None:
                with torch.autocast("cuda", dtype=self.autocast_dtype):
                    heads = self.model(x)
            else:
                heads = self.model(x)
        # K2a takes element strides, so channels-last heads are consumed in place (no NCHW copy)
        out = []
        for h in heads:
            h = h.float()
            if h.stride(2) != h.shape[3] * h.stride(3):
                h = h.contiguous()
            out.append(h)
        return out

    def _meta_dev(self, plan: LetterboxPlan, cls: int) -> Tuple[np.ndarray, torch.Tensor]:
        key = (id(plan), cls)
        if key not in self._meta:
            m = plan.img_meta(cls)
            self._meta[key] = (m, self.ctx.struct_to_device(m))
        return self._meta[key]

    # -------------------------------------------------------------- whole-frame path
    def detect_device(self, frames_dev: torch.Tensor, graph: Optional[bool] = None, imgsz: Optional[int] = None,
                      hook_repeat: bool = False):
        """frames_dev uint8[n,H,W,3] on the GPU -> device tensors (xyxy[n,max_det,4], conf, cls, count)."""
        n, h, w, _ = frames_dev.shape
        if self.head_hook is not None:
            self.head_hook.begin_chunk(n, repeat=hook_repeat)
        imgsz = imgsz or self.imgsz
        plan = self.plan(n, h, w, _ffi.LB_WHOLE, imgsz)
        if self.cuda_graph if graph is None else graph:
            # everything the captured K1 plan / K2a launch bakes in is part of the key
            key = (n, h, w, imgsz, self.conf, self.iou, self.max_det, self.agnostic, id(self.head_hook))
            if key not in self._graphs:
                from .runtime import GraphedStep

                def step(f):
                    xyxy, cf, cl, cnt, (heads, _, _) = self._detect_eager(f, plan)
                    flat = tuple(heads.box) + tuple(heads.cls) if isinstance(heads, SplitHeads) else tuple(heads)
                    return (xyxy, cf, cl, cnt) + flat
                self._graphs[key] = GraphedStep(self.ctx, step, [frames_dev])
            out = self._graphs[key](frames_dev)          # static outputs: valid until the next call with this shape
            meta_h, meta_d = self._meta_dev(plan, 0)
            heads = SplitHeads(out[4:7], out[7:10]) if len(out) == 10 else list(out[4:])
            return out[0], out[1], out[2], out[3], (heads, meta_h, meta_d)
        return self._detect_eager(frames_dev, plan)

    def _detect_eager(self, frames_dev: torch.Tensor, plan: LetterboxPlan):
        n = frames_dev.shape[0]
        with nvtx("hvb:K1a letterbox"):
            x = plan.class_views(plan.run(frames_dev))[0]
        with nvtx("hvb:yolo forward (cuDNN convs + K5)"):
            heads = self.forward_heads(x)
        meta_h, meta_d = self._meta_dev(plan, 0)
        with nvtx("hvb:K2a decode+nms"):
            return self._decode(heads, meta_h, meta_d, n)

    def _decode(self, heads, meta_h, meta_d, n_slots, out=None):
        # meta stays resident on the device; the host copy is only used by the overflow retry
        ctx = self.ctx
        with ctx.lock:
            ctx._enter()
            if out is None:
                out = (ctx.empty((n_slots, self.max_det, 4), torch.float32), ctx.empty((n_slots, self.max_det), torch.float32),
                       ctx.empty((n_slots, self.max_det), torch.int32), torch.zeros((n_slots,), dtype=torch.int32, device=ctx.device))
            xyxy, cf, cl, cnt = out
            ctx.decode_nms_call(heads, self.nc, self.conf, self.iou, self.max_det, self.agnostic, meta_d, xyxy, cf, cl, cnt)
        return xyxy, cf, cl, cnt, (heads, meta_h, meta_d)

    def _retry_overflow(self, xyxy, cf, cl, cnt, state, cnt_host: np.ndarray):
        """Images that reported -1 (> 1024 candidates) are re-run with the 8192-candidate tier."""
        heads, meta_h, meta_d = state
        ctx = self.ctx
        bad = np.nonzero(cnt_host[meta_h["out_slot"]] < 0)[0].astype(np.int32)
        if len(bad) == 0:
            return cnt_host
        with ctx.lock:
            ctx._enter()
            bad_dev = ctx.to_device(bad)
            ctx.decode_nms_call(heads, self.nc, self.conf, self.iou, self.max_det, self.agnostic, meta_d, xyxy, cf, cl, cnt,
                                images_dev=bad_dev, n_images=int(len(bad)))
        cnt_host = cnt.cpu().numpy()
        if (cnt_host[meta_h["out_slot"]] < 0).any():
            raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "more than 8192 candidates above conf=%g in one image" % self.conf)
        return cnt_host

    def detect_batch(self, frames, imgsz: Optional[int] = None) -> List[Detections]:
        """List of Detections (one per frame), equal to from_ultralytics(model(frame, imgsz=imgsz)[0]) per frame."""
        return self.detect_batch_device(self.upload(frames), imgsz)

    def detect_batch_device(self, frames_dev: torch.Tensor, imgsz: Optional[int] = None) -> List[Detections]:
        """detect_batch for frames that are already on the device (uint8[n,H,W,3])."""
        xyxy, cf, cl, cnt, state = self.detect_device(frames_dev, imgsz=imgsz)
        cnt_h = cnt.cpu().numpy()
        if (cnt_h < 0).any():
            cnt_h = self._retry_overflow(xyxy, cf, cl, cnt, state, cnt_h)
        xyxy_h, cf_h, cl_h = xyxy.cpu().numpy(), cf.cpu().numpy(), cl.cpu().numpy()
        return [self._to_detections(xyxy_h[i, :k], cf_h[i, :k], cl_h[i, :k]) for i, k in enumerate(cnt_h)]

    def _to_detections(self, xyxy, conf, cls) -> Detections:
        cls = cls.astype(int)
        names = np.array([self.class_names.get(int(c), str(int(c))) for c in cls], dtype=object) if len(cls) else np.array([], dtype=object)
        return Detections(xyxy=xyxy.astype(np.float32).reshape(-1, 4).copy(), confidence=conf.astype(np.float32).copy(),
                          class_id=cls, tracker_id=None, data={"class_name": names})

    def __call__(self, frame: np.ndarray) -> Detections:
        return self.detect_batch(frame)[0]

    def detect_players(self, frame) -> Detections:
        """VideoProcessor.detect_players: detection + the class / confidence mask of main.py:189-193.  `frame`: a host
        frame uint8[H,W,3], or the same frame already on the device as uint8[1,H,W,3]."""
        d = self.detect_batch_device(frame)[0] if isinstance(frame, torch.Tensor) else self(frame)
        keep = ((d.class_id == PLAYER_CLASS_ID) | (d.class_id == GOALKEEPER_CLASS_ID)) & (d.confidence > self.conf)
        return d[keep]

    # -------------------------------------------------------------- per-tile callback (compat path)
    def as_callback(self, imgsz: int = 640):
        """callback(image_slice) -> Detections for sv.InferenceSlicer(callback=...)."""
        det = self

        def callback(tile: np.ndarray) -> Detections:
            return det.detect_batch(np.ascontiguousarray(tile), imgsz=imgsz)[0]
        return callback
