"""Thin object layer over the C ABI: one `Context` per (process, GPU).

PyTorch is used here only as the device-memory / stream plumbing (tensors own the buffers whose
``data_ptr()`` is handed to libhvb, and work is enqueued on torch's current stream so it orders
with the backbone forwards).  All arithmetic on the hot path happens inside libhvb kernels.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ffi
from ._ffi import check, ptr

class nvtx:
    """NVTX range (shows up in nsys / ncu timelines): `with nvtx("hvb:K1a"): ...`.  A no-op cost of ~1 us."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


_contexts = {}
_contexts_lock = threading.Lock()


def get_context(device: int | str | torch.device | None = None) -> "Context":
    """Process-wide context cache keyed by CUDA device index."""
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    elif isinstance(device, int):
        idx = device
    else:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _ffi.HvbError(_ffi.HVB_ERR_NO_DEVICE, "hvb runs on CUDA devices only (got %r); there is no CPU fallback" % (device,))
        idx = dev.index if dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    with _contexts_lock:
        if idx not in _contexts:
            _contexts[idx] = Context(idx)
        return _contexts[idx]


class SplitHeads:
    """Raw Detect outputs with the class logits in their own dense tensors: box[i] is [B, 64, H_i, W_i], cls[i] is
    [B, nc, H_i, W_i] (any strides with jointly contiguous spatial dims; channels-last is what the K5 runner writes)."""

    def __init__(self, box: Sequence[torch.Tensor], cls: Sequence[torch.Tensor]):
        self.box, self.cls = list(box), list(cls)

    def __len__(self):
        return 3

    def tensors(self) -> List[torch.Tensor]:
        return self.box + self.cls

    def __iter__(self):
        """Combined [B, 64+nc, H, W] tensors (for code that wants the plain Detect output)."""
        return iter(torch.cat((b, c), 1) for b, c in zip(self.box, self.cls))


class Context:
    def __init__(self, device: int = 0):
        self.lib = _ffi.lib()
        self.device_index = device
        h = C.c_void_p()
        check(self.lib.hvb_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = torch.device("cuda", device)
        self.lock = threading.RLock()
        n = C.c_int()
        check(self.lib.hvb_ctx_sm_count(self.handle, C.byref(n)))
        self.sm_count = n.value

    # ------------------------------------------------------------------ plumbing
    def _enter(self):
        """Bind libhvb's stream to torch's current stream on this device."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.hvb_ctx_set_stream(self.handle, C.c_void_p(s)))

    def synchronize(self):
        with self.lock:
            self._enter()
            check(self.lib.hvb_ctx_synchronize(self.handle))

    def retain_buffers(self, on: bool = True) -> None:
        """Keep work buffers alive across growth (call before capturing libhvb launches into a CUDA graph)."""
        check(self.lib.hvb_ctx_retain_buffers(self.handle, 1 if on else 0))

    def launch_count(self, reset: bool = False) -> int:
        n = C.c_uint64()
        check(self.lib.hvb_ctx_launch_count(self.handle, 1 if reset else 0, C.byref(n)))
        return int(n.value)

    def empty(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def to_device(self, arr: np.ndarray, non_blocking: bool = False) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device, non_blocking=non_blocking)

    def struct_to_device(self, arr: np.ndarray) -> torch.Tensor:
        """Upload a numpy structured array as raw bytes."""
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        return torch.from_numpy(raw.copy()).to(self.device)

    # ------------------------------------------------------------------ K1
    def letterbox_plan(self, n_frames: int, frame_h: int, frame_w: int, mode: int = _ffi.LB_WHOLE, imgsz: int = 640,
                       auto: bool = True, stride: int = 32, slice_wh: Tuple[int, int] = (640, 640),
                       overlap_wh: Tuple[int, int] = (128, 128)) -> "LetterboxPlan":
        return LetterboxPlan(self, n_frames, frame_h, frame_w, mode, imgsz, auto, stride, slice_wh, overlap_wh)

    # ------------------------------------------------------------------ K2a
    @staticmethod
    def _level_args(levels: Sequence[torch.Tensor]):
        assert len(levels) == 3
        ptrs = (C.c_void_p * 3)(*[lv.data_ptr() for lv in levels])
        hs = (C.c_int32 * 3)(*[lv.shape[2] for lv in levels])
        ws = (C.c_int32 * 3)(*[lv.shape[3] for lv in levels])
        bs = (C.c_int64 * 3)(*[lv.stride(0) for lv in levels])
        cs = (C.c_int64 * 3)(*[lv.stride(1) for lv in levels])
        for lv in levels:
            if lv.dtype != torch.float32:
                raise TypeError("head tensors must be float32")
            # anchors are row-major over (H, W): need stride(2) == W * stride(3)
            if lv.stride(2) != lv.shape[3] * lv.stride(3):
                raise ValueError("head tensor spatial dims must be jointly contiguous")
        as_ = (C.c_int64 * 3)(*[lv.stride(3) for lv in levels])
        return ptrs, hs, ws, bs, cs, as_

    def decode_nms_call(self, heads, nc, conf, iou, max_det, agnostic, meta_dev, xyxy, cf, cl, cnt, images_dev=None, n_images=None):
        """One K2a launch for either head layout (list of combined tensors, or SplitHeads).  Caller holds the lock."""
        if isinstance(heads, SplitHeads):
            bp, hs, ws, bb, bc, ba = self._level_args(heads.box)
            cp, _, _, cb, cc, ca = self._level_args(heads.cls)
            batch = heads.box[0].shape[0] if images_dev is None else n_images
            check(self.lib.hvb_decode_nms_split(self.handle, bp, cp, hs, ws, bb, bc, ba, cb, cc, ca, ptr(images_dev), batch, nc,
                                                conf, iou, max_det, int(agnostic), ptr(meta_dev), ptr(xyxy), ptr(cf), ptr(cl), ptr(cnt)))
            return
        args = self._level_args(heads)
        if images_dev is None:
            check(self.lib.hvb_decode_nms(self.handle, *args, heads[0].shape[0], nc, conf, iou, max_det, int(agnostic),
                                          ptr(meta_dev), ptr(xyxy), ptr(cf), ptr(cl), ptr(cnt)))
        else:
            check(self.lib.hvb_decode_nms_large(self.handle, *args, ptr(images_dev), n_images, nc, conf, iou, max_det,
                                                int(agnostic), ptr(meta_dev), ptr(xyxy), ptr(cf), ptr(cl), ptr(cnt)))

    def decode_nms(self, levels: Sequence[torch.Tensor], nc: int, conf: float, iou: float = 0.7, max_det: int = 300,
                   agnostic: bool = False, meta: Optional[np.ndarray] = None, n_slots: Optional[int] = None,
                   out: Optional[Tuple[torch.Tensor, ...]] = None, check_overflow: bool = True):
        """Returns (xyxy f32[n_slots,max_det,4], conf f32[n_slots,max_det], cls i32[n_slots,max_det], count i32[n_slots])."""
        B = levels[0].shape[0]
        if meta is None:
            raise ValueError("meta (IMG_META per image) is required")
        n_slots = n_slots if n_slots is not None else B
        with self.lock:
            self._enter()
            meta_dev = meta if isinstance(meta, torch.Tensor) else self.struct_to_device(meta)
            if out is None:
                out = (self.empty((n_slots, max_det, 4), torch.float32), self.empty((n_slots, max_det), torch.float32),
                       self.empty((n_slots, max_det), torch.int32), torch.zeros((n_slots,), dtype=torch.int32, device=self.device))
            xyxy, cf, cl, cnt = out
            args = self._level_args(levels)
            check(self.lib.hvb_decode_nms(self.handle, *args, B, nc, conf, iou, max_det, int(agnostic), ptr(meta_dev),
                                          ptr(xyxy), ptr(cf), ptr(cl), ptr(cnt)))
            if check_overflow:
                # images with more than 1024 candidates report -1: re-run those with the large tier
                slots = meta["out_slot"] if isinstance(meta, np.ndarray) else None
                cnt_h = cnt.cpu().numpy()
                if (cnt_h < 0).any():
                    if slots is None:
                        raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "candidate overflow and meta not available on host")
                    bad = np.nonzero(cnt_h[slots] < 0)[0].astype(np.int32)
                    bad_dev = self.to_device(bad)
                    check(self.lib.hvb_decode_nms_large(self.handle, *args, ptr(bad_dev), int(len(bad)), nc, conf, iou,
                                                        max_det, int(agnostic), ptr(meta_dev), ptr(xyxy), ptr(cf), ptr(cl),
                                                        ptr(cnt)))
                    if (cnt.cpu().numpy() < 0).any():
                        cap = C.c_int()
                        self.lib.hvb_nms_capacity(C.byref(cap))
                        raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY,
                                            "more than %d candidates above conf in one image" % cap.value)
        return xyxy, cf, cl, cnt

    def decode_only(self, levels: Sequence[torch.Tensor], nc: int) -> torch.Tensor:
        B = levels[0].shape[0]
        A = sum(lv.shape[2] * lv.shape[3] for lv in levels)
        with self.lock:
            self._enter()
            out = self.empty((B, 4 + nc, A), torch.float32)
            check(self.lib.hvb_decode_only(self.handle, *self._level_args(levels), B, nc, ptr(out)))
        return out

    def nms_f32(self, boxes: torch.Tensor, scores: torch.Tensor, cls: Optional[torch.Tensor], iou: float,
                max_det: int = 300, agnostic: bool = False) -> torch.Tensor:
        n = boxes.shape[0]
        with self.lock:
            self._enter()
            keep = self.empty((max_det,), torch.int32)
            cnt = torch.zeros((1,), dtype=torch.int32, device=self.device)
            check(self.lib.hvb_nms_f32(self.handle, ptr(boxes.contiguous()), ptr(scores.contiguous()),
                                       ptr(cls.contiguous()) if cls is not None else ptr(None), n, iou, max_det,
                                       int(agnostic or cls is None), ptr(keep), ptr(cnt)))
            k = int(cnt.item())
        return keep[:k]

    # ------------------------------------------------------------------ K2b
    def merge_nms(self, xyxy: torch.Tensor, conf: torch.Tensor, cls: Optional[torch.Tensor], seg_offsets: torch.Tensor,
                  n_segments: int, n_total: int, iou: float, class_agnostic: bool = False) -> torch.Tensor:
        with self.lock:
            self._enter()
            keep = torch.zeros((max(n_total, 1),), dtype=torch.uint8, device=self.device)
            check(self.lib.hvb_merge_nms(self.handle, ptr(xyxy), ptr(conf), ptr(cls), ptr(seg_offsets), n_segments, n_total,
                                         float(iou), int(class_agnostic), ptr(keep)))
        return keep[:n_total]

    def gather_tiles(self, xyxy, conf, cls, count, slot_off_xy: torch.Tensor, slots_per_frame: int, max_det: int):
        n_slots = count.shape[0]
        n_frames = n_slots // slots_per_frame
        with self.lock:
            self._enter()
            cap = n_slots * max_det
            o_xyxy = self.empty((cap, 4), torch.float64)
            o_conf = self.empty((cap,), torch.float32)
            o_cls = self.empty((cap,), torch.int32)
            o_slot = self.empty((cap,), torch.int32)
            seg = self.empty((n_frames + 1,), torch.int32)
            check(self.lib.hvb_gather_tiles(self.handle, ptr(xyxy), ptr(conf), ptr(cls), ptr(count), ptr(slot_off_xy), n_slots,
                                            slots_per_frame, max_det, ptr(o_xyxy), ptr(o_conf), ptr(o_cls), ptr(o_slot), ptr(seg)))
        return o_xyxy, o_conf, o_cls, o_slot, seg

    # ------------------------------------------------------------------ K3
    def crops_from_boxes(self, xyxy: torch.Tensor, frame_idx: Optional[torch.Tensor], frame_h: int, frame_w: int) -> torch.Tensor:
        n = xyxy.shape[0]
        with self.lock:
            self._enter()
            out = self.empty((max(n, 1) * _ffi.CROP_DESC.itemsize,), torch.uint8)
            check(self.lib.hvb_crops_from_boxes(self.handle, ptr(xyxy.contiguous()), ptr(frame_idx), n, frame_h, frame_w, ptr(out)))
        return out

    def color_features(self, pixels: torch.Tensor, crops: torch.Tensor, n: int, roi_mode: int = _ffi.ROI_HYBRID,
                       want_raw: bool = False, out_feat: Optional[torch.Tensor] = None, feat_stride: int = 49):
        with self.lock:
            self._enter()
            if out_feat is None:
                out_feat = self.empty((n, 49), torch.float64)
                feat_stride = 49
            raw = self.empty((max(n, 1) * _ffi.COLOR_RAW.itemsize,), torch.uint8) if want_raw else None
            check(self.lib.hvb_color_features(self.handle, ptr(pixels), ptr(crops), n, roi_mode, ptr(out_feat), feat_stride, ptr(raw)))
        return (out_feat, raw) if want_raw else out_feat

    def jersey_color_stats(self, pixels: torch.Tensor, crops: torch.Tensor, n: int, roi_mode: int = _ffi.ROI_SEGMENT):
        """hvb_jersey_raw[n] as a device uint8 tensor (view it with ``_ffi.JERSEY_RAW`` after ``.cpu().numpy()``)."""
        with self.lock:
            self._enter()
            raw = self.empty((max(n, 1) * _ffi.JERSEY_RAW.itemsize,), torch.uint8)
            check(self.lib.hvb_jersey_color_stats(self.handle, ptr(pixels), ptr(crops), n, roi_mode, ptr(raw)))
        return raw

    def cvt_hsv_lab(self, bgr: torch.Tensor):
        n_px = bgr.numel() // 3
        with self.lock:
            self._enter()
            hsv = torch.empty_like(bgr)
            lab = torch.empty_like(bgr)
            check(self.lib.hvb_cvt_hsv_lab(self.handle, ptr(bgr), n_px, ptr(hsv), ptr(lab)))
        return hsv, lab

    def mnv3_preprocess(self, pixels: torch.Tensor, crops: torch.Tensor, n: int, roi_mode: int = _ffi.ROI_HYBRID,
                        want_u8: bool = False, rows: Optional[int] = None, out: Optional[torch.Tensor] = None):
        """`rows` >= n: allocate that many output rows (the kernel fills the first n, the rest are zero) so that the
        backbone behind it can run at a bucketed batch size.  `out`: write into this float32[rows >= n, 3, 128, 64]
        tensor instead (the static input of a captured trunk graph); its rows beyond n are zeroed."""
        with self.lock:
            self._enter()
            if out is not None:
                assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape[1:]) == (3, 128, 64) and out.shape[0] >= n
                if out.shape[0] > n:
                    out[n:].zero_()
            elif rows is not None and rows > n:
                out = self.empty((rows, 3, 128, 64), torch.float32)
                out[n:].zero_()
            else:
                out = self.empty((n, 3, 128, 64), torch.float32)
            u8 = self.empty((n, 128, 64, 3), torch.uint8) if want_u8 else None
            valid = self.empty((max(n, 1),), torch.uint8)
            check(self.lib.hvb_mnv3_preprocess(self.handle, ptr(pixels), ptr(crops), n, roi_mode, ptr(out), ptr(u8), ptr(valid)))
        return (out, valid[:n], u8) if want_u8 else (out, valid[:n])

    # ------------------------------------------------------------------ K4
    def standardize(self, x: torch.Tensor):
        n, d = x.shape
        with self.lock:
            self._enter()
            mean = self.empty((d,), torch.float64)
            scale = self.empty((d,), torch.float64)
            xs = self.empty((n, d), torch.float64)
            check(self.lib.hvb_standardize(self.handle, ptr(x.contiguous()), n, d, ptr(mean), ptr(scale), ptr(xs)))
        return mean, scale, xs

    def scale_transform(self, x: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
        n, d = x.shape
        with self.lock:
            self._enter()
            xs = self.empty((n, d), torch.float64)
            check(self.lib.hvb_scale_transform(self.handle, ptr(x.contiguous()), n, d, ptr(mean), ptr(scale), ptr(xs)))
        return xs

    def gram_affinity(self, x: torch.Tensor, gamma: float = 1.0, mode: int = 0, want_d2: bool = True, want_a: bool = True):
        n, d = x.shape
        with self.lock:
            self._enter()
            d2 = self.empty((n, n), torch.float64) if want_d2 else None
            a = self.empty((n, n), torch.float64) if want_a else None
            check(self.lib.hvb_gram_affinity(self.handle, ptr(x.contiguous()), n, d, float(gamma), mode, ptr(d2), ptr(a)))
        return d2, a

    def gram_tc(self, x: torch.Tensor) -> torch.Tensor:
        n, d = x.shape
        with self.lock:
            self._enter()
            g = self.empty((n, n), torch.float32)
            check(self.lib.hvb_gram_tc(self.handle, ptr(x.contiguous()), n, d, ptr(g)))
        return g

    def iou_cost(self, a: torch.Tensor, b: torch.Tensor, scores: Optional[torch.Tensor], a_off: torch.Tensor,
                 b_off: torch.Tensor, out_off: torch.Tensor, n_problems: int, max_na: int, max_nb: int, out_size: int,
                 flags: int = 0):
        with self.lock:
            self._enter()
            out = self.empty((max(out_size, 1),), torch.float64)
            check(self.lib.hvb_iou_cost(self.handle, ptr(a), ptr(b), ptr(scores), ptr(a_off), ptr(b_off), ptr(out_off),
                                        n_problems, max_na, max_nb, flags, ptr(out)))
        return out

    # ------------------------------------------------------------------ K7 (device ByteTrack)
    def bytetrack_create(self, n_clips: int, track_activation_threshold: float, det_threshold: float,
                         minimum_matching_threshold: float, max_time_lost: int, minimum_consecutive_frames: int) -> C.c_void_p:
        h = C.c_void_p()
        with self.lock:
            self._enter()
            check(self.lib.hvb_bytetrack_create(self.handle, n_clips, float(track_activation_threshold), float(det_threshold),
                                                float(minimum_matching_threshold), int(max_time_lost),
                                                int(minimum_consecutive_frames), C.byref(h)))
        return h

    def bytetrack_destroy(self, handle) -> None:
        with self.lock:
            check(self.lib.hvb_bytetrack_destroy(self.handle, handle))

    def bytetrack_reset(self, handle) -> None:
        with self.lock:
            self._enter()
            check(self.lib.hvb_bytetrack_reset(self.handle, handle))

    def bytetrack_update(self, handle, xyxy: torch.Tensor, conf: torch.Tensor, cls: Optional[torch.Tensor], count: torch.Tensor,
                         n_frames: int, clip_stride: int, frame_stride: int = 1, min_conf: float = float("-inf"),
                         class_mask: int = 0xFFFFFFFF, seq: int = 0):
        """K2a-layout detections [images, max_det, ...] -> (row i32[images,max_det], tracker_id i32[images,max_det],
        count i32[images]) on the device; one launch, stream-ordered."""
        images, max_det = conf.shape
        with self.lock:
            self._enter()
            row = self.empty((images, max_det), torch.int32)
            tid = self.empty((images, max_det), torch.int32)
            cnt = torch.zeros((images,), dtype=torch.int32, device=self.device)
            check(self.lib.hvb_bytetrack_update(self.handle, handle, ptr(xyxy), ptr(conf), ptr(cls), ptr(count), n_frames, max_det,
                                                clip_stride, frame_stride, float(min_conf), class_mask & 0xFFFFFFFF, int(seq),
                                                ptr(row), ptr(tid), ptr(cnt)))
        return row, tid, cnt

    # ------------------------------------------------------------------ K8 (device spectral clustering pieces)
    def laplacian_normalize(self, a: torch.Tensor):
        """float64[n,n] affinity -> (M = D^-1/2 A0 D^-1/2 float64[n,n], dd = sqrt(degree) float64[n])."""
        n = a.shape[0]
        a = a.contiguous()
        with self.lock:
            self._enter()
            m = self.empty((n, n), torch.float64)
            dd = self.empty((n,), torch.float64)
            check(self.lib.hvb_laplacian_normalize(self.handle, ptr(a), n, ptr(m), ptr(dd)))
        return m, dd

    def sym_block_matvec(self, m: torch.Tensor, x: torch.Tensor, shift: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Y = (M + shift I) X for x float64[8, n] (each vector contiguous)."""
        n = m.shape[0]
        with self.lock:
            self._enter()
            y = out if out is not None else self.empty((8, n), torch.float64)
            check(self.lib.hvb_sym_block_matvec(self.handle, ptr(m), n, ptr(x), float(shift), ptr(y)))
        return y

    def block_gram(self, a: torch.Tensor, b: torch.Tensor, mode: int = 0) -> torch.Tensor:
        """mode 0: float64[65] with [:64] = A^T B (8 x 8 row-major); mode 1: [:64] = R^-1 of A^T A = R^T R, [64] = not-PD flag."""
        with self.lock:
            self._enter()
            out = self.empty((65,), torch.float64)
            check(self.lib.hvb_block_gram(self.handle, ptr(a), ptr(b), a.shape[1], mode, ptr(out)))
        return out

    def block_rotate(self, x: torch.Tensor, y: Optional[torch.Tensor], q: torch.Tensor, lam: Optional[torch.Tensor] = None):
        """In place: X <- X Q, Y <- Y Q (q float64[8,8] on the device); with lam float64[8]: returns |Y q_j - lam_j X q_j|^2."""
        with self.lock:
            self._enter()
            res = self.empty((8,), torch.float64) if lam is not None else None
            check(self.lib.hvb_block_rotate(self.handle, ptr(x), ptr(y), x.shape[1], ptr(q), ptr(lam), ptr(res)))
        return res

    def kmeans_lloyd(self, x: torch.Tensor, init_centers: torch.Tensor, max_iter: int, tol: float):
        """x float64[n,d] (mean-centred), init_centers float64[n_init,k,d] -> labels i32[n_init,n], centres, inertia,
        n_iter, flags — every initialisation in one launch."""
        n, d = x.shape
        n_init, k, _ = init_centers.shape
        x, init_centers = x.contiguous(), init_centers.contiguous()
        with self.lock:
            self._enter()
            labels = self.empty((n_init, n), torch.int32)
            centers = self.empty((n_init, k, d), torch.float64)
            inertia = self.empty((n_init,), torch.float64)
            n_iter = self.empty((n_init,), torch.int32)
            flags = self.empty((n_init,), torch.int32)
            check(self.lib.hvb_kmeans_lloyd(self.handle, ptr(x), n, d, k, ptr(init_centers), n_init, int(max_iter), float(tol),
                                            ptr(labels), ptr(centers), ptr(inertia), ptr(n_iter), ptr(flags)))
        return labels, centers, inertia, n_iter, flags

    # ------------------------------------------------------------------ feature exchange (NCCL bound inside libhvb)
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        check(self.lib.hvb_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_create(self, unique_id: bytes, world: int, rank: int) -> C.c_void_p:
        """Collective over all ranks: ncclCommInitRank on this context's device."""
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        with self.lock:
            check(self.lib.hvb_comm_create(self.handle, C.cast(buf, C.c_void_p), int(world), int(rank), C.byref(h)))
        return h

    def comm_destroy(self, comm) -> None:
        with self.lock:
            check(self.lib.hvb_comm_destroy(self.handle, comm))

    def allgather_features(self, comm, local: torch.Tensor, world: int) -> Tuple[torch.Tensor, np.ndarray]:
        """local float64[n_g, d] on this device -> (float64[sum n_g, d] in rank order, the per-rank row counts)."""
        if local.dtype != torch.float64 or local.dim() != 2:
            raise ValueError("expected a float64 [n, d] tensor")
        local = local.contiguous()
        n, d = local.shape
        counts = np.zeros((world,), np.int32)
        total = C.c_int64()
        with self.lock:
            self._enter()
            check(self.lib.hvb_allgather_counts(self.handle, comm, n, ptr(counts), C.byref(total)))
            out = self.empty((int(total.value), d), torch.float64)
            if total.value:
                check(self.lib.hvb_allgather_features(self.handle, comm, ptr(local) if n else None, n, d, ptr(counts), ptr(out)))
        return out, counts

    # ------------------------------------------------------------------ K5 (backbone glue; NHWC float32)
    ACT = {"none": 0, "silu": 1, "relu": 2, "hardswish": 3, "silu_fast": 4}

    @staticmethod
    def _nhwc(t: torch.Tensor) -> Tuple[int, int]:
        """(pixels, channels) of a [N,C,H,W] tensor stored channels-last (dense NHWC)."""
        if t.dtype != torch.float32 or t.dim() != 4 or not t.is_contiguous(memory_format=torch.channels_last):
            raise ValueError("expected a float32 channels_last [N,C,H,W] tensor, got %s %s" % (t.dtype, tuple(t.stride())))
        return t.shape[0] * t.shape[2] * t.shape[3], t.shape[1]

    def bias_act(self, x: torch.Tensor, bias: Optional[torch.Tensor], act: str = "silu", residual: Optional[torch.Tensor] = None,
                 out1: Optional[torch.Tensor] = None, out1_off: int = 0, out2: Optional[torch.Tensor] = None,
                 out2_off: int = 0, c2_begin: int = 0, c2_count: Optional[int] = None, write_out1: bool = True,
                 out2_upsample2: bool = False) -> torch.Tensor:
        """y = act(x + bias) (+ residual) in one pass; y -> out1[:, out1_off:out1_off+C] (default: in place on x)
        and, for channels [c2_begin, c2_begin+c2_count), -> out2[:, out2_off:...].  Returns out1 (or out2)."""
        npix, c = self._nhwc(x)
        if out1 is None and write_out1:
            out1 = x
        for t in (residual, out1, out2):
            if t is not None and self._nhwc(t)[0] != (4 * npix if (t is out2 and out2_upsample2) else npix):
                raise ValueError("pixel count mismatch")
        uh, uw = (x.shape[2], x.shape[3]) if (out2 is not None and out2_upsample2) else (0, 0)
        if residual is not None and residual.shape[1] != c:
            raise ValueError("residual channel mismatch")
        c2 = (c if c2_count is None else c2_count) if out2 is not None else 0
        with self.lock:
            self._enter()
            check(self.lib.hvb_bias_act(self.handle, ptr(x), ptr(bias), ptr(residual), npix, c, self.ACT[act],
                                        ptr(out1), out1.shape[1] if out1 is not None else 0, out1_off,
                                        ptr(out2), out2.shape[1] if out2 is not None else 0, out2_off, c2_begin, c2, uh, uw))
        return out1 if out1 is not None else out2

    def pointwise_conv(self, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act: str = "silu_fast",
                       out1: Optional[torch.Tensor] = None, out1_off: int = 0, out2: Optional[torch.Tensor] = None,
                       out2_off: int = 0, c2_begin: int = 0, c2_count: Optional[int] = None) -> torch.Tensor:
        """K6: 1x1 convolution (weight [c_out, c_in] or [c_out, c_in, 1, 1]) + bias + activation of a channels-last
        tensor in one tcgen05 GEMM; destinations as in bias_act.  Raises HvbError(UNSUPPORTED) for channel counts the
        kernel does not tile (the caller then keeps conv2d + bias_act)."""
        npix, cin = self._nhwc(x)
        cout = weight.shape[0]
        if weight.numel() != cout * cin or not weight.is_contiguous() and not weight.is_contiguous(memory_format=torch.channels_last):
            raise ValueError("weight must be a dense [c_out, c_in(,1,1)] tensor")
        if out1 is None:
            out1 = torch.empty((x.shape[0], cout, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device,
                               memory_format=torch.channels_last)
        for t in (out1, out2):
            if t is not None and self._nhwc(t)[0] != npix:
                raise ValueError("pixel count mismatch")
        c2 = (cout if c2_count is None else c2_count) if out2 is not None else 0
        with self.lock:
            self._enter()
            check(self.lib.hvb_pointwise_conv(self.handle, ptr(x), cin, ptr(weight), ptr(bias), npix, cin, cout, self.ACT[act],
                                              ptr(out1), out1.shape[1], out1_off, ptr(out2),
                                              out2.shape[1] if out2 is not None else 0, out2_off, c2_begin, c2))
        return out1

    def concat_nhwc(self, sources: Sequence[torch.Tensor], shifts: Optional[Sequence[int]] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Channel concat of up to 4 channels-last tensors; source i is nearest-upsampled by 2**shifts[i]."""
        shifts = list(shifts) if shifts is not None else [0] * len(sources)
        n = sources[0].shape[0]
        h, w = sources[0].shape[2] << shifts[0], sources[0].shape[3] << shifts[0]
        for t, s in zip(sources, shifts):
            self._nhwc(t)
            if (t.shape[0], t.shape[2] << s, t.shape[3] << s) != (n, h, w):
                raise ValueError("concat sources disagree on the output size")
        ctot = sum(t.shape[1] for t in sources)
        with self.lock:
            self._enter()
            if out is None:
                out = torch.empty((n, ctot, h, w), dtype=torch.float32, device=self.device, memory_format=torch.channels_last)
            k = len(sources)
            srcs = (C.c_void_p * 4)(*([t.data_ptr() for t in sources] + [0] * (4 - k)))
            chans = (C.c_int32 * 4)(*([t.shape[1] for t in sources] + [0] * (4 - k)))
            shs = (C.c_int32 * 4)(*(shifts + [0] * (4 - k)))
            check(self.lib.hvb_concat_nhwc(self.handle, srcs, chans, shs, k, n, h, w, ptr(out)))
        return out

    def sppf_pool_concat(self, y0: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[y0, m(y0), m(m(y0)), m(m(m(y0)))] along channels, m = MaxPool2d(5, 1, 2); channels-last in and out."""
        self._nhwc(y0)
        n, c, h, w = y0.shape
        with self.lock:
            self._enter()
            if out is None:
                out = torch.empty((n, 4 * c, h, w), dtype=torch.float32, device=self.device, memory_format=torch.channels_last)
            check(self.lib.hvb_sppf_pool_concat(self.handle, ptr(y0), n, h, w, c, ptr(out)))
        return out

    def stem_conv(self, x_nchw: torch.Tensor, weight_host: np.ndarray, bias_host: Optional[np.ndarray]) -> torch.Tensor:
        """Layer 0 (3->C, 3x3, stride 2, pad 1) + bias + SiLU: NCHW float32 in, channels-last out."""
        n, ci, h, w = x_nchw.shape
        if ci != 3 or x_nchw.dtype != torch.float32 or not x_nchw.is_contiguous():
            raise ValueError("stem_conv expects a dense float32 [N,3,H,W] tensor")
        co = int(weight_host.shape[0])
        assert weight_host.dtype == np.float32 and weight_host.shape == (co, 3, 3, 3) and weight_host.flags.c_contiguous
        with self.lock:
            self._enter()
            out = torch.empty((n, co, (h - 1) // 2 + 1, (w - 1) // 2 + 1), dtype=torch.float32, device=self.device,
                              memory_format=torch.channels_last)
            check(self.lib.hvb_stem_conv(self.handle, ptr(x_nchw), ptr(weight_host), ptr(bias_host), n, h, w, co, ptr(out)))
        return out

    # ------------------------------------------------------------------ host-buffer entry points
    def color_features_host(self, pixels: np.ndarray, crops: np.ndarray, roi_mode: int = _ffi.ROI_HYBRID, want_raw: bool = False):
        n = len(crops)
        feat = np.empty((n, 49), np.float64)
        raw = np.zeros((n,), _ffi.COLOR_RAW) if want_raw else None
        if n:
            with self.lock:
                self._enter()
                check(self.lib.hvb_color_features_host(self.handle, ptr(pixels), pixels.nbytes, ptr(crops), n, roi_mode,
                                                       ptr(feat), ptr(raw)))
        return (feat, raw) if want_raw else feat

    def jersey_color_stats_host(self, pixels: np.ndarray, crops: np.ndarray, roi_mode: int = _ffi.ROI_SEGMENT) -> np.ndarray:
        n = len(crops)
        raw = np.zeros((n,), _ffi.JERSEY_RAW)
        if n:
            with self.lock:
                self._enter()
                check(self.lib.hvb_jersey_color_stats_host(self.handle, ptr(pixels), pixels.nbytes, ptr(crops), n, roi_mode, ptr(raw)))
        return raw

    def mnv3_preprocess_host(self, pixels: np.ndarray, crops: np.ndarray, roi_mode: int = _ffi.ROI_HYBRID):
        n = len(crops)
        out = np.empty((n, 3, 128, 64), np.float32)
        valid = np.zeros((n,), np.uint8)
        if n:
            with self.lock:
                self._enter()
                check(self.lib.hvb_mnv3_preprocess_host(self.handle, ptr(pixels), pixels.nbytes, ptr(crops), n, roi_mode,
                                                        ptr(out), ptr(valid)))
        return out, valid

    def merge_nms_host(self, xyxy: np.ndarray, conf: np.ndarray, cls: Optional[np.ndarray], iou: float,
                       class_agnostic: bool = False) -> np.ndarray:
        n = len(xyxy)
        keep = np.zeros((n,), np.uint8)
        if n:
            xyxy = np.ascontiguousarray(xyxy, np.float64)
            conf = np.ascontiguousarray(conf, np.float32)
            cls32 = np.ascontiguousarray(cls, np.int32) if cls is not None else None
            with self.lock:
                self._enter()
                check(self.lib.hvb_merge_nms_host(self.handle, ptr(xyxy), ptr(conf), ptr(cls32), n, float(iou),
                                                  int(class_agnostic), ptr(keep)))
            if (keep == 0xFF).any():
                raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "merge NMS segment exceeds the on-chip capacity")
        return keep.astype(bool)

    def iou_cost_host(self, a: np.ndarray, b: np.ndarray, scores: Optional[np.ndarray] = None) -> np.ndarray:
        a, b = np.asarray(a), np.asarray(b)
        flags = (1 if a.dtype == np.float32 else 0) | (2 if b.dtype == np.float32 else 0)
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 4)
        b = np.ascontiguousarray(b, np.float64).reshape(-1, 4)
        out = np.zeros((len(a), len(b)), np.float64)
        if len(a) and len(b):
            sc = np.ascontiguousarray(scores, np.float64) if scores is not None else None
            with self.lock:
                self._enter()
                check(self.lib.hvb_iou_cost_host(self.handle, ptr(a), len(a), ptr(b), len(b), ptr(sc), flags, ptr(out)))
        return out

    def gram_affinity_host(self, x: np.ndarray, gamma: float = 1.0, mode: int = 0):
        x = np.ascontiguousarray(x, np.float64)
        n, d = x.shape
        d2 = np.empty((n, n), np.float64)
        a = np.empty((n, n), np.float64)
        with self.lock:
            self._enter()
            check(self.lib.hvb_gram_affinity_host(self.handle, ptr(x), n, d, float(gamma), mode, ptr(d2), ptr(a)))
        return d2, a


class LetterboxPlan:
    """hvb_lb_plan: geometry for a chunk of equally sized frames; `run` is one kernel launch."""

    def __init__(self, ctx: Context, n_frames, frame_h, frame_w, mode, imgsz, auto, stride, slice_wh, overlap_wh):
        self.ctx = ctx
        self.n_frames, self.frame_h, self.frame_w, self.mode, self.imgsz = n_frames, frame_h, frame_w, mode, imgsz
        h = C.c_void_p()
        with ctx.lock:
            check(ctx.lib.hvb_lb_plan_create(ctx.handle, n_frames, frame_h, frame_w, mode, imgsz, int(auto), stride,
                                             slice_wh[0], slice_wh[1], overlap_wh[0], overlap_wh[1], C.byref(h)))
        self.handle = h
        n = C.c_int()
        check(ctx.lib.hvb_lb_plan_num_classes(h, C.byref(n)))
        self.classes = np.zeros((n.value,), _ffi.LB_CLASS)
        for i in range(n.value):
            check(ctx.lib.hvb_lb_plan_get_class(h, i, ptr(self.classes[i:i + 1])))
        check(ctx.lib.hvb_lb_plan_num_tiles(h, C.byref(n)))
        self.tiles = np.zeros((n.value,), _ffi.LB_TILE)
        check(ctx.lib.hvb_lb_plan_get_tiles(h, ptr(self.tiles), n.value))
        self.tiles_per_frame = n.value // n_frames
        f = C.c_int64()
        check(ctx.lib.hvb_lb_plan_out_floats(h, C.byref(f)))
        self.out_floats = int(f.value)
        r, w = C.c_int64(), C.c_int64()
        check(ctx.lib.hvb_lb_plan_bytes(h, C.byref(r), C.byref(w)))
        self.read_bytes, self.write_bytes = int(r.value), int(w.value)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.hvb_lb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def class_views(self, out: torch.Tensor) -> List[torch.Tensor]:
        """Views of the flat output buffer as one [batch,3,out_h,out_w] tensor per shape class."""
        views = []
        for c in self.classes:
            numel = int(c["batch"]) * 3 * int(c["out_h"]) * int(c["out_w"])
            o = int(c["out_offset"])
            views.append(out[o:o + numel].view(int(c["batch"]), 3, int(c["out_h"]), int(c["out_w"])))
        return views

    def run(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8[n_frames, H, W, 3] on the context's GPU -> flat float32 buffer."""
        assert frames.dtype == torch.uint8 and frames.is_contiguous()
        assert tuple(frames.shape) == (self.n_frames, self.frame_h, self.frame_w, 3), frames.shape
        with self.ctx.lock:
            self.ctx._enter()
            if out is None:
                out = self.ctx.empty((self.out_floats,), torch.float32)
            check(self.ctx.lib.hvb_lb_plan_run(self.handle, ptr(frames), ptr(out)))
        return out

    def run_u8(self, frames: torch.Tensor) -> torch.Tensor:
        assert frames.dtype == torch.uint8 and frames.is_contiguous()
        with self.ctx.lock:
            self.ctx._enter()
            out = self.ctx.empty((self.out_floats,), torch.uint8)
            check(self.ctx.lib.hvb_lb_plan_run_u8(self.handle, ptr(frames), ptr(out)))
        return out

    def img_meta(self, cls: int, slot_of_tile=None) -> np.ndarray:
        """IMG_META rows for the batch of shape class `cls`, in batch order.  out_slot defaults to
        frame * tiles_per_frame + tile (slicer order), the layout hvb_gather_tiles expects."""
        t = self.tiles[self.tiles["cls"] == cls]
        t = t[np.argsort(t["batch_index"], kind="stable")]
        m = np.zeros((len(t),), _ffi.IMG_META)
        m["gain"], m["pad_x"], m["pad_y"] = t["gain"], t["pad_x"], t["pad_y"]
        m["clip_w"], m["clip_h"] = t["src_w"], t["src_h"]
        m["off_x"], m["off_y"] = t["src_x"], t["src_y"]
        m["out_slot"] = t["frame"] * self.tiles_per_frame + t["tile"]
        return m

    def slot_offsets(self) -> np.ndarray:
        """float32[n_slots, 2] tile offsets in slot order (frame-major, slicer tile order)."""
        return np.stack([self.tiles["src_x"], self.tiles["src_y"]], 1).astype(np.float32)


class GraphedStep:
    """A fixed-shape device step captured once into a CUDA graph and replayed.

    The hot path of one chunk is several hundred launches (library convolutions + libhvb kernels); for the small
    batches of the sliced path the host cannot issue them as fast as the GPU retires them.  `fn(*tensors)` must be
    free of host synchronisation and must only depend on its tensor arguments' CONTENTS (shapes fixed).  The step is
    run eagerly first (cuDNN autotuning, libhvb plan / work-buffer allocation), then captured; later calls copy the
    inputs into the graph's OWN input buffers (a device-to-device copy, ~0.1 ms per 400 MB) and replay, so callers may
    recycle their buffers freely.  The returned tensors are the graph's static outputs: consume them before the next
    call."""

    def __init__(self, ctx: Context, fn, example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        self.ctx, self.fn = ctx, fn
        self.static_in = [t.clone() for t in example_inputs]
        self._launches = 0
        ctx.retain_buffers(True)
        side = torch.cuda.Stream(device=ctx.device)
        side.wait_stream(torch.cuda.current_stream(ctx.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream(ctx.device).wait_stream(side)
        torch.cuda.synchronize(ctx.device)
        before = ctx.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        self._launches = ctx.launch_count() - before      # libhvb kernels inside one replay

    @property
    def launches_per_replay(self) -> int:
        return self._launches

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
