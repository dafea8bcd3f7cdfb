"""CPU restatement of the `supervision` pieces on the reference's hot path.
TEST INFRASTRUCTURE — see oracle/__init__.py.

PARITY UNPINNED: supervision is a third-party dependency, not vendored under /root/reference and
not installed in this image (version unpinned by the reference; >= 0.21 is required by
``sv.ByteTrack(minimum_consecutive_frames=...)`` at hockey/main.py:167).  Restated from its
published algorithm:

  crop_image                 supervision/utils/image.py::crop_image        (hockey/main.py:326)
  generate_offsets           supervision/detection/tools/inference_slicer.py::InferenceSlicer._generate_offset
  move_boxes                 supervision/detection/utils.py::move_boxes
  box_iou_batch              supervision/detection/utils.py::box_iou_batch
  box_non_max_suppression    supervision/detection/utils.py::box_non_max_suppression
  with_nms                   supervision/detection/core.py::Detections.with_nms
  run_slicer                 InferenceSlicer.__call__ (thread_workers=1: deterministic tile order)
  iou_distance / fuse_score  supervision/tracker/byte_tracker/matching.py     (hockey/main.py:228,265)
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def crop_image(image: np.ndarray, xyxy) -> np.ndarray:
    """np.round (half-to-even) then a numpy slice: returns a VIEW, numpy slice semantics apply."""
    xyxy = np.round(np.asarray(xyxy)).astype(int)
    x0, y0, x1, y1 = xyxy.flatten()
    return image[y0:y1, x0:x1]


def generate_offsets(resolution_wh: Tuple[int, int], slice_wh: Tuple[int, int] = (640, 640),
                     overlap_ratio_wh: Optional[Tuple[float, float]] = (0.2, 0.2),
                     overlap_wh: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """int[n,4] (xmin,ymin,xmax,ymax), row-major over y then x; edge tiles clipped, not shifted."""
    slice_w, slice_h = slice_wh
    img_w, img_h = resolution_wh
    if overlap_wh is None:
        ow = int(overlap_ratio_wh[0] * slice_w)
        oh = int(overlap_ratio_wh[1] * slice_h)
    else:
        ow, oh = overlap_wh
    ws = np.arange(0, img_w, slice_w - ow)
    hs = np.arange(0, img_h, slice_h - oh)
    xmin, ymin = np.meshgrid(ws, hs)
    xmax = np.clip(xmin + slice_w, 0, img_w)
    ymax = np.clip(ymin + slice_h, 0, img_h)
    return np.stack([xmin, ymin, xmax, ymax], axis=-1).reshape(-1, 4)


def move_boxes(xyxy: np.ndarray, offset) -> np.ndarray:
    """xyxy + [ox, oy, ox, oy]; integer offsets upcast float32 boxes to float64 like numpy does."""
    offset = np.asarray(offset)
    return xyxy + np.hstack([offset, offset])


def box_iou_batch(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    def area(x):
        return (x[2] - x[0]) * (x[3] - x[1])
    area_a = area(a.T)
    area_b = area(b.T)
    tl = np.maximum(a[:, None, :2], b[:, :2])
    br = np.minimum(a[:, None, 2:], b[:, 2:])
    inter = np.prod(np.clip(br - tl, a_min=0, a_max=None), 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area_a[:, None] + area_b - inter)
    return np.nan_to_num(iou)


def box_non_max_suppression(predictions: np.ndarray, iou_threshold: float = 0.5) -> np.ndarray:
    """predictions float64[N,5|6] (x1,y1,x2,y2,score[,category]) -> bool[N] keep mask in INPUT order."""
    rows, columns = predictions.shape
    if rows == 0:
        return np.zeros(0, bool)
    if columns == 5:
        predictions = np.c_[predictions, np.zeros(rows)]
    sort_index = np.flip(predictions[:, 4].argsort())
    predictions = predictions[sort_index]
    boxes = predictions[:, :4]
    categories = predictions[:, 5]
    ious = box_iou_batch(boxes, boxes)
    ious = ious - np.eye(rows)
    keep = np.ones(rows, dtype=bool)
    for index, (iou, category) in enumerate(zip(ious, categories)):
        if not keep[index]:
            continue
        condition = (iou > iou_threshold) & (categories == category)
        keep = keep & ~condition
    return keep[sort_index.argsort()]


def with_nms(xyxy: np.ndarray, confidence: np.ndarray, class_id: Optional[np.ndarray],
             threshold: float = 0.5, class_agnostic: bool = False) -> np.ndarray:
    """Detections.with_nms: hstack -> float64, returns the boolean keep mask."""
    if len(xyxy) == 0:
        return np.zeros(0, bool)
    if class_agnostic or class_id is None:
        pred = np.hstack((xyxy, confidence.reshape(-1, 1)))
    else:
        pred = np.hstack((xyxy, confidence.reshape(-1, 1), class_id.reshape(-1, 1)))
    return box_non_max_suppression(pred.astype(np.float64), threshold)


def run_slicer(image: np.ndarray, callback: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]],
               slice_wh=(640, 640), overlap_ratio_wh=(0.2, 0.2), overlap_wh=None,
               iou_threshold: float = 0.5, overlap_filter: str = "nms", class_agnostic: bool = False):
    """InferenceSlicer.__call__ with thread_workers=1.  `callback(tile)` returns (xyxy f32[n,4],
    conf f32[n], cls int[n]).  Returns merged (xyxy float64[m,4], conf, cls) after the filter."""
    h, w = image.shape[:2]
    offsets = generate_offsets((w, h), slice_wh, overlap_ratio_wh, overlap_wh)
    X, C, K = [], [], []
    for off in offsets:
        tile = crop_image(image, off)
        xyxy, conf, cls = callback(tile)
        X.append(move_boxes(np.asarray(xyxy).reshape(-1, 4), off[:2]).astype(np.float64))
        C.append(np.asarray(conf, np.float32).reshape(-1))
        K.append(np.asarray(cls).astype(np.int64).reshape(-1))
    xyxy = np.concatenate(X) if X else np.zeros((0, 4))
    conf = np.concatenate(C) if C else np.zeros(0, np.float32)
    cls = np.concatenate(K) if K else np.zeros(0, np.int64)
    if overlap_filter == "none" or len(xyxy) == 0:
        return xyxy, conf, cls
    keep = with_nms(xyxy, conf, cls, iou_threshold, class_agnostic)
    return xyxy[keep], conf[keep], cls[keep]


# ------------------------------------------------------------------ ByteTrack matching costs
def iou_distance(atlbrs: np.ndarray, btlbrs: np.ndarray) -> np.ndarray:
    """1 - box_iou_batch(np.asarray(atlbrs), np.asarray(btlbrs)) (matching.iou_distance).  The dtypes
    are NOT unified first: detection boxes are float32 (their area is then computed in float32 by
    numpy before promotion), Kalman track boxes float64."""
    a = np.asarray(atlbrs).reshape(-1, 4)
    b = np.asarray(btlbrs).reshape(-1, 4)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((a.shape[0], b.shape[0]), dtype=float)
    return 1 - box_iou_batch(a, b)


def fuse_score(cost_matrix: np.ndarray, det_scores: np.ndarray) -> np.ndarray:
    """1 - (1 - cost) * score  (matching.fuse_score)."""
    if cost_matrix.size == 0:
        return cost_matrix
    iou_sim = 1 - cost_matrix
    det_scores = np.expand_dims(np.asarray(det_scores, dtype=float), axis=0).repeat(cost_matrix.shape[0], axis=0)
    return 1 - iou_sim * det_scores
