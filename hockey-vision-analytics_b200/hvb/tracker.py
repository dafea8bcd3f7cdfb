"""ByteTrack — drop-in for ``sv.ByteTrack`` as the reference uses it (hockey/main.py:162-168,
207-211 construct it; :228 and :265 call ``update_with_detections``).  SURVEY.md §8(f) rank 1.

The association logic (three matching rounds, lost/removed bookkeeping, id issuing) is sequential
per clip and runs on the host, exactly like the reference; what moves to the GPU is the arithmetic
that grows with tracks x detections: every ``iou_distance`` / ``fuse_score`` matrix comes from the
K4b kernel (``hvb_iou_cost``), bit-identical to numpy's ``box_iou_batch``.  State is kept as
structure-of-arrays (one row per track) and the xyah Kalman filter is evaluated for all tracks of a
round at once.  Hungarian assignment is scipy's, as in supervision.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.optimize import linear_sum_assignment

from .detections import Detections

_NEW, _TRACKED, _LOST, _REMOVED = 0, 1, 2, 3
_W_POS, _W_VEL = 1.0 / 20, 1.0 / 160
_F = np.eye(8)
_F[:4, 4:] = np.eye(4)
_H = np.eye(4, 8)


def _assign(cost: np.ndarray, thresh: float):
    """matching.linear_assignment: clamp, Hungarian, keep matches with cost <= thresh."""
    na, nb = cost.shape
    if cost.size == 0:
        return np.empty((0, 2), dtype=int), list(range(na)), list(range(nb))
    c = np.where(cost > thresh, thresh + 1e-4, cost)
    rows, cols = linear_sum_assignment(c)
    ok = c[rows, cols] <= thresh
    m = np.column_stack((rows[ok], cols[ok]))
    ua = sorted(set(range(na)) - set(m[:, 0].tolist()))
    ub = sorted(set(range(nb)) - set(m[:, 1].tolist()))
    return m, ua, ub


class ByteTrack:
    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: int = 30, minimum_consecutive_frames: int = 1,
                 iou_cost: Optional[Callable] = None, device="cuda:0"):
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.det_thresh = track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self._device = device
        self._iou_cost = iou_cost          # (a_boxes, b_boxes, scores|None) -> cost matrix; default = K4b kernel
        self.reset()

    # ------------------------------------------------------------------ state
    def reset(self):
        self.frame_id = 0
        self._next_internal, self._next_external = 0, 1
        cap = 64
        self._mean = np.zeros((cap, 8))
        self._cov = np.zeros((cap, 8, 8))
        self._state = np.zeros(cap, np.int8)
        self._activated = np.zeros(cap, bool)
        self._frame = np.zeros(cap, np.int64)
        self._start = np.zeros(cap, np.int64)
        self._len = np.zeros(cap, np.int64)
        self._score = np.zeros(cap)
        self._ext = np.full(cap, -1, np.int64)
        self._n = 0                        # rows in use == internal ids issued
        self.tracked: List[int] = []       # ordered like supervision's lists (order decides Hungarian ties)
        self.lost: List[int] = []
        self.removed: List[int] = []

    def _grow(self):
        for name in ("_mean", "_cov", "_state", "_activated", "_frame", "_start", "_len", "_score", "_ext"):
            a = getattr(self, name)
            b = np.zeros((2 * a.shape[0],) + a.shape[1:], a.dtype)
            if name == "_ext":
                b[:] = -1
            b[: a.shape[0]] = a
            setattr(self, name, b)

    def _cost(self, a: np.ndarray, b: np.ndarray, scores: Optional[np.ndarray] = None) -> np.ndarray:
        if len(a) == 0 or len(b) == 0:
            return np.zeros((len(a), len(b)))
        if self._iou_cost is None:
            from .runtime import get_context
            ctx = get_context(self._device)
            self._iou_cost = lambda x, y, s=None: ctx.iou_cost_host(x, y, s)
        return self._iou_cost(a, b, scores)

    def _tlbr(self, idx: Sequence[int]) -> np.ndarray:
        """Track boxes (float64) from the Kalman means: xyah -> tlwh -> tlbr."""
        m = self._mean[list(idx), :4].copy()
        m[:, 2] *= m[:, 3]
        m[:, :2] -= m[:, 2:] / 2
        m[:, 2:] += m[:, :2]
        return m

    # ------------------------------------------------------------------ Kalman filter (batched)
    @staticmethod
    def _xyah(tlwh: np.ndarray) -> np.ndarray:
        r = np.asarray(tlwh).copy()
        r[..., :2] += r[..., 2:] / 2
        r[..., 2] /= r[..., 3]
        return r

    def _predict(self, idx: List[int]):
        if not idx:
            return
        idx = np.asarray(idx)
        mean = self._mean[idx].copy()
        mean[self._state[idx] != _TRACKED, 7] = 0
        h = mean[:, 3]
        std = np.stack([_W_POS * h, _W_POS * h, np.full_like(h, 1e-2), _W_POS * h,
                        _W_VEL * h, _W_VEL * h, np.full_like(h, 1e-5), _W_VEL * h], 1)
        q = np.zeros((len(idx), 8, 8))
        q[:, np.arange(8), np.arange(8)] = np.square(std)
        self._mean[idx] = mean @ _F.T
        self._cov[idx] = _F @ self._cov[idx] @ _F.T + q

    def _correct(self, i: int, xyah: np.ndarray):
        mean, cov = self._mean[i], self._cov[i]
        h = mean[3]
        r = np.diag(np.square([_W_POS * h, _W_POS * h, 1e-1, _W_POS * h]))
        pm = _H @ mean
        pc = _H @ cov @ _H.T + r
        gain = np.linalg.solve(pc, (cov @ _H.T).T).T
        self._mean[i] = mean + (xyah - pm) @ gain.T
        self._cov[i] = cov - gain @ pc @ gain.T

    def _initiate(self, tlwh: np.ndarray, score: float) -> int:
        if self._n == len(self._state):
            self._grow()
        i = self._n
        self._n += 1
        self._next_internal += 1
        z = self._xyah(np.asarray(tlwh, np.float32))
        h = z[3]
        std = [2 * _W_POS * h, 2 * _W_POS * h, 1e-2, 2 * _W_POS * h, 10 * _W_VEL * h, 10 * _W_VEL * h, 1e-5, 10 * _W_VEL * h]
        self._mean[i] = np.r_[z, np.zeros(4)]
        self._cov[i] = np.diag(np.square(std))
        self._state[i], self._len[i], self._score[i] = _TRACKED, 0, score
        self._activated[i] = self.frame_id == 1
        self._ext[i] = -1
        if self.minimum_consecutive_frames == 1:
            self._ext[i] = self._next_external
            self._next_external += 1
        self._frame[i] = self._start[i] = self.frame_id
        return i

    def _hit(self, i: int, tlwh32: np.ndarray, score: float, reactivate: bool):
        self._correct(i, self._xyah(tlwh32))
        self._state[i], self._frame[i], self._score[i] = _TRACKED, self.frame_id, score
        if reactivate:
            self._len[i] = 0
        else:
            self._len[i] += 1
            if self._len[i] == self.minimum_consecutive_frames:
                self._activated[i] = True
                if self._ext[i] == -1:
                    self._ext[i] = self._next_external
                    self._next_external += 1

    # ------------------------------------------------------------------ one frame
    def _step(self, xyxy: np.ndarray, scores: np.ndarray) -> List[int]:
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        hi = scores > self.track_activation_threshold
        lo = (scores > 0.1) & (scores < self.track_activation_threshold)
        det_xyxy, det_s = xyxy[hi].astype(np.float32), scores[hi]
        det_tlwh = det_xyxy.copy()
        det_tlwh[:, 2:] -= det_tlwh[:, :2]
        det_box = det_tlwh.copy()
        det_box[:, 2:] += det_box[:, :2]                    # float32 tlbr exactly as STrack.tlbr gives it

        unconfirmed = [t for t in self.tracked if not self._activated[t]]
        confirmed = [t for t in self.tracked if self._activated[t]]
        pool = confirmed + [t for t in self.lost if t not in set(confirmed)]
        self._predict(pool)

        # round 1: all confirmed + lost tracks vs high-score detections, IoU fused with the score
        m, u_trk, u_det = _assign(self._cost(self._tlbr(pool), det_box, det_s.astype(np.float64)), self.minimum_matching_threshold)
        for it, idt in m:
            t = pool[it]
            was_tracked = self._state[t] == _TRACKED
            self._hit(t, det_tlwh[idt], det_s[idt], reactivate=not was_tracked)
            (activated if was_tracked else refind).append(t)

        # round 2: still-tracked leftovers vs low-score detections
        lo_xyxy, lo_s = xyxy[lo].astype(np.float32), scores[lo]
        lo_tlwh = lo_xyxy.copy()
        lo_tlwh[:, 2:] -= lo_tlwh[:, :2]
        lo_box = lo_tlwh.copy()
        lo_box[:, 2:] += lo_box[:, :2]
        rest = [pool[i] for i in u_trk if self._state[pool[i]] == _TRACKED]
        m2, u_rest, _ = _assign(self._cost(self._tlbr(rest), lo_box), 0.5)
        for it, idt in m2:
            t = rest[it]
            self._hit(t, lo_tlwh[idt], lo_s[idt], reactivate=False)
            activated.append(t)
        for it in u_rest:
            t = rest[it]
            if self._state[t] != _LOST:
                self._state[t] = _LOST
                lost.append(t)

        # round 3: unconfirmed tracks vs the remaining high-score detections
        rem = list(u_det)
        m3, u_unc, u_rem = _assign(self._cost(self._tlbr(unconfirmed), det_box[rem], det_s[rem].astype(np.float64)), 0.7)
        for it, idt in m3:
            t = unconfirmed[it]
            self._hit(t, det_tlwh[rem[idt]], det_s[rem[idt]], reactivate=False)
            activated.append(t)
        for it in u_unc:
            self._state[unconfirmed[it]] = _REMOVED
            removed.append(unconfirmed[it])

        for k in u_rem:                                     # births
            j = rem[k]
            if det_s[j] < self.det_thresh:
                continue
            activated.append(self._initiate(det_tlwh[j], det_s[j]))

        for t in self.lost:
            if self.frame_id - self._frame[t] > self.max_time_lost:
                self._state[t] = _REMOVED
                removed.append(t)

        def joint(a, b):
            seen = set(a)
            return a + [t for t in b if not (t in seen or seen.add(t))]

        self.tracked = [t for t in self.tracked if self._state[t] == _TRACKED]
        self.tracked = joint(joint(self.tracked, activated), refind)
        tr = set(self.tracked)
        self.lost = [t for t in self.lost if t not in tr] + lost
        rm = set(self.removed)
        self.lost = [t for t in self.lost if t not in rm]
        self.removed = removed
        # duplicates between tracked and lost (IoU distance < 0.15): keep the longer-lived one
        if self.tracked and self.lost:
            d = self._cost(self._tlbr(self.tracked), self._tlbr(self.lost))
            da, db = set(), set()
            for ia, ib in zip(*np.where(d < 0.15)):
                ta, tb = self.tracked[ia], self.lost[ib]
                if self._frame[ta] - self._start[ta] > self._frame[tb] - self._start[tb]:
                    db.add(ib)
                else:
                    da.add(ia)
            self.tracked = [t for i, t in enumerate(self.tracked) if i not in da]
            self.lost = [t for i, t in enumerate(self.lost) if i not in db]
        return [t for t in self.tracked if self._activated[t]]

    def update_with_tensors(self, tensors: np.ndarray) -> List[int]:
        t = np.asarray(tensors)
        return self._step(t[:, :4], t[:, 4])

    def update_with_detections(self, detections: Detections) -> Detections:
        xyxy = np.asarray(detections.xyxy, np.float32).reshape(-1, 4)
        conf = np.asarray(detections.confidence, np.float32).reshape(-1) if len(xyxy) else np.zeros(0, np.float32)
        out = self._step(xyxy, conf)
        if len(out) > 0 and len(xyxy) > 0:
            m, _, _ = _assign(self._cost(xyxy, self._tlbr(out)), 0.5)
            ids = np.full(len(xyxy), -1, dtype=int)
            for i_det, i_trk in m:
                ids[i_det] = int(self._ext[out[i_trk]])
            detections.tracker_id = ids
            return detections[ids != -1]
        empty = Detections.empty()
        empty.tracker_id = np.array([], dtype=int)
        return empty
