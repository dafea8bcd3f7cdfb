"""Minimal ``sv.Detections``-compatible container.

The reference exchanges ``supervision.Detections`` objects between its stages
(hockey/main.py:186-193, 225-237, 265-287).  When ``supervision`` is importable the real class is
used (``Detections = sv.Detections``); otherwise this shim provides the subset of its surface the
hot path touches: the six fields, ``len``, boolean / integer / slice indexing, ``empty``,
``merge``, ``is_empty`` and ``with_nms`` (which runs the K2b kernel, never numpy).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Union

import numpy as np

try:  # pragma: no cover - supervision is not installed in the build image
    import supervision as _sv
    HAVE_SUPERVISION = True
except Exception:  # noqa: BLE001
    _sv = None
    HAVE_SUPERVISION = False


@dataclass
class _ShimDetections:
    xyxy: np.ndarray
    mask: Optional[np.ndarray] = None
    confidence: Optional[np.ndarray] = None
    class_id: Optional[np.ndarray] = None
    tracker_id: Optional[np.ndarray] = None
    data: Dict[str, Union[np.ndarray, List]] = field(default_factory=dict)

    def __post_init__(self):
        self.xyxy = np.asarray(self.xyxy).reshape(-1, 4)
        n = len(self.xyxy)
        for name in ("confidence", "class_id", "tracker_id"):
            v = getattr(self, name)
            if v is not None and len(v) != n:
                raise ValueError("%s must have %d entries, got %d" % (name, n, len(v)))

    def __len__(self) -> int:
        return len(self.xyxy)

    def __iter__(self) -> Iterator:
        for i in range(len(self)):
            yield (self.xyxy[i], None if self.mask is None else self.mask[i],
                   None if self.confidence is None else self.confidence[i],
                   None if self.class_id is None else self.class_id[i],
                   None if self.tracker_id is None else self.tracker_id[i],
                   {k: v[i] for k, v in self.data.items()})

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.data.get(index)
        if isinstance(index, int):
            index = [index]
        idx = np.asarray(index) if not isinstance(index, slice) else index

        def take(v):
            if v is None:
                return None
            return np.asarray(v)[idx]

        return type(self)(xyxy=self.xyxy[idx], mask=take(self.mask), confidence=take(self.confidence),
                          class_id=take(self.class_id), tracker_id=take(self.tracker_id),
                          data={k: take(v) for k, v in self.data.items()})

    @classmethod
    def empty(cls):
        return cls(xyxy=np.empty((0, 4), dtype=np.float32), confidence=np.array([], dtype=np.float32),
                   class_id=np.array([], dtype=int))

    def is_empty(self) -> bool:
        return len(self) == 0

    @classmethod
    def merge(cls, detections_list: List["_ShimDetections"]):
        dets = [d for d in detections_list if not d.is_empty()]
        if not dets:
            return cls.empty()
        if len(dets) == 1:
            return dets[0]

        def cat(name):
            vals = [getattr(d, name) for d in dets]
            if all(v is None for v in vals):
                return None
            if any(v is None for v in vals):
                raise ValueError("cannot merge detections with inconsistent `%s`" % name)
            return np.concatenate([np.asarray(v) for v in vals])

        keys = set().union(*[set(d.data.keys()) for d in dets])
        data = {k: np.concatenate([np.asarray(d.data[k]) for d in dets]) for k in keys
                if all(k in d.data for d in dets)}
        return cls(xyxy=np.vstack([d.xyxy for d in dets]), mask=cat("mask"), confidence=cat("confidence"),
                   class_id=cat("class_id"), tracker_id=cat("tracker_id"), data=data)

    def with_nms(self, threshold: float = 0.5, class_agnostic: bool = False):
        """Detections.with_nms through the K2b kernel (float64, keep mask in input order)."""
        if len(self) == 0:
            return self
        if self.confidence is None:
            raise AssertionError("Detections confidence must be given for NMS to be executed.")
        from .runtime import get_context
        keep = get_context().merge_nms_host(self.xyxy, self.confidence,
                                            None if (class_agnostic or self.class_id is None) else self.class_id,
                                            threshold, class_agnostic)
        return self[keep]


Detections = _sv.Detections if HAVE_SUPERVISION else _ShimDetections


def crop_image(image: np.ndarray, xyxy) -> np.ndarray:
    """sv.crop_image (hockey/main.py:326): np.round -> int, numpy slice view."""
    xyxy = np.round(np.asarray(xyxy)).astype(int)
    x0, y0, x1, y1 = xyxy.flatten()
    return image[y0:y1, x0:x1]
