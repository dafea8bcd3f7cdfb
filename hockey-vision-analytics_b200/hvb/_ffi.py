"""ctypes binding of libhvb.so (C ABI: include/hvb.h).

There is no CPU fallback: importing this module without a built ``libhvb.so`` raises, and
creating a context without a B200 raises ``HvbError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HVB_LIB") or os.path.join(_HERE, "libhvb.so")     # HVB_LIB: A/B kernel variants


class HvbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("libhvb status %d: %s" % (status, message))
        self.status = status


HVB_ERR_CUDA, HVB_ERR_ARG, HVB_ERR_NO_DEVICE, HVB_ERR_CAPACITY, HVB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
LB_WHOLE, LB_SLICE_EXACT, LB_SLICE_UNIFORM = 0, 1, 2
ROI_HYBRID, ROI_SIMPLE, ROI_WHOLE, ROI_SEGMENT = 0, 1, 2, 3

# numpy mirrors of the C structs -------------------------------------------------------------
LB_CLASS = np.dtype([("out_h", "<i4"), ("out_w", "<i4"), ("tiles_per_frame", "<i4"), ("batch", "<i4"),
                     ("out_offset", "<i8")], align=True)
LB_TILE = np.dtype([("frame", "<i4"), ("tile", "<i4"), ("cls", "<i4"), ("batch_index", "<i4"),
                    ("src_x", "<i4"), ("src_y", "<i4"), ("src_w", "<i4"), ("src_h", "<i4"),
                    ("new_w", "<i4"), ("new_h", "<i4"), ("top", "<i4"), ("left", "<i4"),
                    ("out_h", "<i4"), ("out_w", "<i4"), ("gain", "<f4"), ("pad_x", "<f4"), ("pad_y", "<f4")],
                   align=True)
IMG_META = np.dtype([("gain", "<f4"), ("pad_x", "<f4"), ("pad_y", "<f4"), ("clip_w", "<f4"), ("clip_h", "<f4"),
                     ("off_x", "<f4"), ("off_y", "<f4"), ("out_slot", "<i4")], align=True)
CROP_DESC = np.dtype([("offset", "<i8"), ("pitch", "<i4"), ("h", "<i4"), ("w", "<i4"), ("reserved", "<i4")],
                     align=True)
COLOR_RAW = np.dtype([("hist", "<u4", (34,)), ("counts", "<u4", (3,)), ("n", "<u4"), ("roi", "<u4", (4,)),
                      ("pad_", "<u4", (2,)), ("sums", "<u8", (6,)), ("sumsq", "<u8", (6,))], align=True)
assert LB_CLASS.itemsize == 24 and LB_TILE.itemsize == 68 and IMG_META.itemsize == 32
JERSEY_RAW = np.dtype([("n", "<u4"), ("white", "<u4"), ("hue_hist", "<u4", (18,)), ("sat_colored", "<u8"),
                       ("sat_all", "<u8"), ("val_all", "<u8"), ("roi", "<i4", (4,))], align=True)
assert CROP_DESC.itemsize == 24 and COLOR_RAW.itemsize == 272 and JERSEY_RAW.itemsize == 120

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_pp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is int unless noted); also the list the symbol-export test checks.
PROTOTYPES = {
    "hvb_version": [],
    "hvb_last_error": [],
    "hvb_device_count": [C.POINTER(_i)],
    "hvb_ctx_create": [_i, _pp],
    "hvb_ctx_destroy": [_vp],
    "hvb_ctx_set_stream": [_vp, _vp],
    "hvb_ctx_use_own_stream": [_vp],
    "hvb_ctx_get_stream": [_vp, _pp],
    "hvb_ctx_synchronize": [_vp],
    "hvb_ctx_retain_buffers": [_vp, _i],
    "hvb_ctx_sm_count": [_vp, C.POINTER(_i)],
    "hvb_malloc": [_vp, _sz, _pp],
    "hvb_free": [_vp, _vp],
    "hvb_host_alloc": [_vp, _sz, _pp],
    "hvb_host_free": [_vp, _vp],
    "hvb_memcpy_h2d": [_vp, _vp, _vp, _sz],
    "hvb_memcpy_d2h": [_vp, _vp, _vp, _sz],
    "hvb_memset": [_vp, _vp, _i, _sz],
    "hvb_stage_frames": [_vp, _i, _sz, _vp, _i],
    "hvb_ctx_launch_count": [_vp, _i, C.POINTER(C.c_uint64)],
    "hvb_timer_start": [_vp],
    "hvb_timer_stop_ms": [_vp, C.POINTER(_f)],
    "hvb_lb_plan_create": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _pp],
    "hvb_lb_plan_destroy": [_vp],
    "hvb_lb_plan_num_classes": [_vp, C.POINTER(_i)],
    "hvb_lb_plan_get_class": [_vp, _i, _vp],
    "hvb_lb_plan_num_tiles": [_vp, C.POINTER(_i)],
    "hvb_lb_plan_get_tiles": [_vp, _vp, _i],
    "hvb_lb_plan_out_floats": [_vp, C.POINTER(_i64)],
    "hvb_lb_plan_bytes": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "hvb_lb_plan_run": [_vp, _vp, _vp],
    "hvb_lb_plan_run_u8": [_vp, _vp, _vp],
    "hvb_decode_nms": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "hvb_decode_nms_large": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "hvb_decode_nms_split": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "hvb_decode_only": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "hvb_nms_f32": [_vp, _vp, _vp, _vp, _i, _f, _i, _i, _vp, _vp],
    "hvb_nms_capacity": [C.POINTER(_i)],
    "hvb_merge_nms": [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _i, _vp],
    "hvb_gather_tiles": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "hvb_crops_from_boxes": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "hvb_color_features": [_vp, _vp, _vp, _i, _i, _vp, _i64, _vp],
    "hvb_cvt_hsv_lab": [_vp, _vp, _i64, _vp, _vp],
    "hvb_mnv3_preprocess": [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "hvb_standardize": [_vp, _vp, _i, _i, _vp, _vp, _vp],
    "hvb_scale_transform": [_vp, _vp, _i, _i, _vp, _vp, _vp],
    "hvb_gram_affinity": [_vp, _vp, _i, _i, _d, _i, _vp, _vp],
    "hvb_gram_tc": [_vp, _vp, _i, _i, _vp],
    "hvb_iou_cost": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "hvb_bias_act": [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i],
    "hvb_concat_nhwc": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "hvb_sppf_pool_concat": [_vp, _vp, _i, _i, _i, _i, _vp],
    "hvb_stem_conv": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "hvb_pointwise_conv": [_vp, _vp, _i, _vp, _vp, _i64, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _i],
    "hvb_bytetrack_create": [_vp, _i, _d, _d, _d, _i, _i, _pp],
    "hvb_bytetrack_destroy": [_vp, _vp],
    "hvb_bytetrack_reset": [_vp, _vp],
    "hvb_bytetrack_update": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _i64, _f, C.c_uint32, _i, _vp, _vp, _vp],
    "hvb_spectral_block": [C.POINTER(_i)],
    "hvb_laplacian_normalize": [_vp, _vp, _i, _vp, _vp],
    "hvb_sym_block_matvec": [_vp, _vp, _i, _vp, _d, _vp],
    "hvb_block_gram": [_vp, _vp, _vp, _i, _i, _vp],
    "hvb_block_rotate": [_vp, _vp, _vp, _i, _vp, _vp, _vp],
    "hvb_kmeans_lloyd": [_vp, _vp, _i, _i, _i, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _vp],
    "hvb_comm_unique_id": [_vp],
    "hvb_comm_create": [_vp, _vp, _i, _i, _pp],
    "hvb_comm_wrap": [_vp, _vp, _pp],
    "hvb_comm_destroy": [_vp, _vp],
    "hvb_comm_info": [_vp, C.POINTER(_i), C.POINTER(_i)],
    "hvb_allgather_counts": [_vp, _vp, _i, _vp, C.POINTER(_i64)],
    "hvb_allgather_features": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "hvb_color_features_host": [_vp, _vp, _sz, _vp, _i, _i, _vp, _vp],
    "hvb_jersey_color_stats": [_vp, _vp, _vp, _i, _i, _vp],
    "hvb_jersey_color_stats_host": [_vp, _vp, _sz, _vp, _i, _i, _vp],
    "hvb_mnv3_preprocess_host": [_vp, _vp, _sz, _vp, _i, _i, _vp, _vp],
    "hvb_merge_nms_host": [_vp, _vp, _vp, _vp, _i, _d, _i, _vp],
    "hvb_iou_cost_host": [_vp, _vp, _i, _vp, _i, _vp, _i, _vp],
    "hvb_gram_affinity_host": [_vp, _vp, _i, _i, _d, _i, _vp, _vp],
}

_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load libhvb.so (once).  Raises if it has not been built — there is no fallback path."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    "libhvb.so not found at %s — build it with `python hockey-vision-analytics_b200/build.py` "
                    "(or __graft_entry__.build()); hvb has no CPU fallback" % LIB_PATH)
            L = C.CDLL(LIB_PATH)
            for name, args in PROTOTYPES.items():
                fn = getattr(L, name)
                fn.argtypes = args
                fn.restype = C.c_char_p if name == "hvb_last_error" else C.c_int
            _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise HvbError(status, lib().hvb_last_error().decode("utf-8", "replace"))


def ptr(x) -> C.c_void_p:
    """Device/host pointer of a torch tensor, numpy array, int or None as c_void_p."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, C.c_void_p):
        return x
    raise TypeError("cannot take a pointer of %r" % type(x))
