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

    def _correct_many(self, idx: Sequence[int], xyah: np.ndarray):
        """KalmanFilter.update for several tracks at once (each track is hit at most once per association round, so
        the per-track updates of a round are independent): one batched 4x4 solve instead of a Python loop of them."""
        idx = list(idx)
        if not idx:
            return
        mean, cov = self._mean[idx], self._cov[idx]                       # [M,8], [M,8,8]
        h = mean[:, 3]
        r = np.zeros((len(idx), 4, 4))
        sp = np.square(_W_POS * h)
        r[:, 0, 0] = r[:, 1, 1] = r[:, 3, 3] = sp
        r[:, 2, 2] = 1e-2
        pm = mean[:, :4]                                                  # H = [I 0]
        pc = cov[:, :4, :4] + r
        b = cov[:, :, :4]                                                 # cov @ H.T
        gain = np.linalg.solve(pc, b.transpose(0, 2, 1)).transpose(0, 2, 1)   # [M,8,4]
        self._mean[idx] = mean + np.einsum("mk,mjk->mj", xyah - pm, gain)
        self._cov[idx] = cov - gain @ pc @ gain.transpose(0, 2, 1)

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

    def _hit(self, i: int, tlwh32: np.ndarray, score: float, reactivate: bool, corrected: bool = False):
        if not corrected:
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

    def _hit_many(self, tr: np.ndarray, scores: np.ndarray, was_tracked: np.ndarray):
        """STrack.update / re_activate bookkeeping for the tracks matched in one association round (already corrected by
        _correct_many), vectorised; external ids are still handed out in match order."""
        if len(tr) == 0:
            return
        self._state[tr] = _TRACKED
        self._frame[tr] = self.frame_id
        self._score[tr] = scores
        self._len[tr[~was_tracked]] = 0
        up = tr[was_tracked]
        self._len[up] += 1
        newly = up[self._len[up] == self.minimum_consecutive_frames]
        if len(newly):
            self._activated[newly] = True
            for t in newly:
                if self._ext[t] == -1:
                    self._ext[t] = self._next_external
                    self._next_external += 1

    # ------------------------------------------------------------------ one frame
    def _step(self, xyxy: np.ndarray, scores: np.ndarray) -> List[int]:
        return self._drive(self._step_gen(xyxy, scores))

    def _drive(self, gen):
        """Run a cost-requesting generator to completion with this tracker's own cost function."""
        try:
            req = next(gen)
            while True:
                req = gen.send(self._cost(*req))
        except StopIteration as stop:
            return stop.value

    def _step_gen(self, xyxy: np.ndarray, scores: np.ndarray):
        """BYTETracker.update_with_tensors as a generator: every IoU cost matrix it needs is requested with
        ``cost = yield (a_boxes, b_boxes, scores_or_None)`` so that a driver can batch the requests of many clips into
        one K4b launch (MultiClipByteTrack); the single-clip path answers them one by one (_drive)."""
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        hi = scores > self.track_activation_threshold
        lo = (scores > 0.1) & (scores < self.track_activation_threshold)
        det_xyxy, det_s = xyxy[hi].astype(np.float32), scores[hi]
        det_tlwh = det_xyxy.copy()
        det_tlwh[:, 2:] -= det_tlwh[:, :2]
        det_box = det_tlwh.copy()
        det_box[:, 2:] += det_box[:, :2]                    # float32 tlbr exactly as STrack.tlbr gives it

        act = self._activated
        unconfirmed = [t for t in self.tracked if not act[t]]
        confirmed = [t for t in self.tracked if act[t]]
        cset = set(confirmed)
        pool = confirmed + [t for t in self.lost if t not in cset]
        self._predict(pool)

        # round 1: all confirmed + lost tracks vs high-score detections, IoU fused with the score
        m, u_trk, u_det = _assign((yield (self._tlbr(pool), det_box, det_s.astype(np.float64))), self.minimum_matching_threshold)
        tr1 = np.asarray([pool[it] for it, _ in m], dtype=np.int64)
        di1 = np.asarray([idt for _, idt in m], dtype=np.int64)
        self._correct_many(tr1.tolist(), self._xyah(det_tlwh[di1]))
        was1 = self._state[tr1] == _TRACKED
        self._hit_many(tr1, det_s[di1], was1)
        activated.extend(tr1[was1].tolist())
        refind.extend(tr1[~was1].tolist())

        # round 2: still-tracked leftovers vs low-score detections
        lo_xyxy, lo_s = xyxy[lo].astype(np.float32), scores[lo]
        lo_tlwh = lo_xyxy.copy()
        lo_tlwh[:, 2:] -= lo_tlwh[:, :2]
        lo_box = lo_tlwh.copy()
        lo_box[:, 2:] += lo_box[:, :2]
        rest = [pool[i] for i in u_trk if self._state[pool[i]] == _TRACKED]
        m2, u_rest, _ = _assign((yield (self._tlbr(rest), lo_box, None)), 0.5)
        tr2 = np.asarray([rest[it] for it, _ in m2], dtype=np.int64)
        di2 = np.asarray([idt for _, idt in m2], dtype=np.int64)
        self._correct_many(tr2.tolist(), self._xyah(lo_tlwh[di2]))
        self._hit_many(tr2, lo_s[di2], np.ones(len(tr2), bool))
        activated.extend(tr2.tolist())
        for it in u_rest:
            t = rest[it]
            if self._state[t] != _LOST:
                self._state[t] = _LOST
                lost.append(t)

        # round 3: unconfirmed tracks vs the remaining high-score detections
        rem = list(u_det)
        m3, u_unc, u_rem = _assign((yield (self._tlbr(unconfirmed), det_box[rem], det_s[rem].astype(np.float64))), 0.7)
        tr3 = np.asarray([unconfirmed[it] for it, _ in m3], dtype=np.int64)
        di3 = np.asarray([rem[idt] for _, idt in m3], dtype=np.int64)
        self._correct_many(tr3.tolist(), self._xyah(det_tlwh[di3]))
        self._hit_many(tr3, det_s[di3], np.ones(len(tr3), bool))
        activated.extend(tr3.tolist())
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
            d = yield (self._tlbr(self.tracked), self._tlbr(self.lost), None)
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
        return self._drive(self._update_gen(detections))

    def _update_gen(self, detections: Detections):
        xyxy = np.asarray(detections.xyxy, np.float32).reshape(-1, 4)
        conf = np.asarray(detections.confidence, np.float32).reshape(-1) if len(xyxy) else np.zeros(0, np.float32)
        out = yield from self._step_gen(xyxy, conf)
        if len(out) > 0 and len(xyxy) > 0:
            m, _, _ = _assign((yield (xyxy, self._tlbr(out), None)), 0.5)
            ids = np.full(len(xyxy), -1, dtype=int)
            for i_det, i_trk in m:
                ids[i_det] = int(self._ext[out[i_trk]])
            detections.tracker_id = ids
            return detections[ids != -1]
        empty = Detections.empty()
        empty.tracker_id = np.array([], dtype=int)
        return empty


class DeviceByteTrack:
    """sv.ByteTrack on the device (K7, csrc/k7_bytetrack.cu): same constructor as ``sv.ByteTrack`` (main.py:162-168), same
    ``update_with_detections`` — Kalman filter, the three association rounds, the assignment solver and the id
    bookkeeping all run in one kernel launch per call — plus ``update_chunk_device``, which steps a whole chunk of frames
    of every clip straight from K2a's device outputs without any host round trip.  ``n_clips`` independent trackers share
    one object (one warp each, concurrently).  Results equal hvb.tracker.ByteTrack / the restated supervision tracker."""

    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: int = 30, minimum_consecutive_frames: int = 1,
                 n_clips: int = 1, device="cuda:0"):
        from .runtime import get_context
        self.ctx = get_context(device)
        self.n_clips = n_clips
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.det_thresh = track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self._h = self.ctx.bytetrack_create(n_clips, track_activation_threshold, self.det_thresh, minimum_matching_threshold,
                                            self.max_time_lost, minimum_consecutive_frames)
        self._seq = 0                       # sequence number of the next chunk (the device accepts chunks strictly in order)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.ctx.bytetrack_destroy(self._h)
                self._h = None
        except Exception:                                  # noqa: BLE001 - interpreter shutdown
            pass

    def reset(self) -> None:
        self.ctx.bytetrack_reset(self._h)
        self._seq = 0

    def next_seq(self) -> int:
        s = self._seq
        self._seq += 1
        return s

    def update_chunk_device(self, xyxy, conf, cls, count, n_frames: Optional[int] = None, min_conf: float = float("-inf"),
                            class_mask: int = 0xFFFFFFFF, seq: Optional[int] = None):
        """K2a outputs of n_frames consecutive frames per clip, clip-major (image = clip * n_frames + frame) ->
        (row, tracker_id, count) device tensors in the same image order.  count -1 / -2: see include/hvb.h.
        `seq`: None = the next chunk; an explicit number resubmits a chunk the device rejected (count -2)."""
        images = conf.shape[0]
        n_frames = images // self.n_clips if n_frames is None else n_frames
        if n_frames * self.n_clips != images:
            raise ValueError("expected n_clips * n_frames = %d images, got %d" % (n_frames * self.n_clips, images))
        return self.ctx.bytetrack_update(self._h, xyxy, conf, cls, count, n_frames, clip_stride=n_frames, frame_stride=1,
                                         min_conf=min_conf, class_mask=class_mask, seq=self.next_seq() if seq is None else seq)

    def update_many(self, detections: Sequence[Detections]) -> List[Detections]:
        """One frame of every clip (detections[i] belongs to clip i), host in / host out, one launch."""
        import torch
        from . import _ffi
        assert len(detections) == self.n_clips
        n = [len(d) for d in detections]
        md = max(max(n), 1)
        xy = np.zeros((self.n_clips, md, 4), np.float32)
        cf = np.zeros((self.n_clips, md), np.float32)
        for i, d in enumerate(detections):
            if n[i]:
                xy[i, :n[i]] = np.asarray(d.xyxy, np.float32).reshape(-1, 4)
                cf[i, :n[i]] = np.asarray(d.confidence, np.float32).reshape(-1)
        dev = self.ctx.device
        row, tid, cnt = self.ctx.bytetrack_update(self._h, torch.from_numpy(xy).to(dev), torch.from_numpy(cf).to(dev), None,
                                                  torch.tensor(n, dtype=torch.int32, device=dev), 1, clip_stride=1, frame_stride=1,
                                                  seq=self.next_seq())
        row_h, tid_h, cnt_h = row.cpu().numpy(), tid.cpu().numpy(), cnt.cpu().numpy()
        if (cnt_h < 0).any():
            raise _ffi.HvbError(_ffi.HVB_ERR_CAPACITY, "ByteTrack capacity exceeded (256 live tracks per clip, 320 detections per frame)")
        out = []
        for i, d in enumerate(detections):
            k = int(cnt_h[i])
            if k == 0:
                e = Detections.empty()
                e.tracker_id = np.array([], dtype=int)
                out.append(e)
                continue
            ids = np.full(n[i], -1, dtype=int)
            ids[row_h[i, :k]] = tid_h[i, :k]
            d.tracker_id = ids
            out.append(d[ids != -1])
        return out

    def update_with_detections(self, detections: Detections) -> Detections:
        if self.n_clips != 1:
            raise ValueError("update_with_detections is the single-clip call; use update_many")
        return self.update_many([detections])[0]


class MultiClipByteTrack:
    """N independent ByteTrack instances (one per clip) stepped in lockstep, one frame of every clip per call.

    Tracking is sequential per clip, but the clips are independent: the IoU cost matrices that the N
    trackers need in the same association round (first / second / unconfirmed association, duplicate removal,
    track -> detection id mapping) are computed by ONE batched K4b launch (``hvb_iou_cost`` takes per-problem offsets)
    and one device -> host copy, instead of one launch + copy per clip and round: 5 launches per frame for any number
    of clips.  Results are those of N separate ``ByteTrack`` objects (tests/test_tracker.py)."""

    def __init__(self, n_clips: int, device="cuda:0", batched_cost: Optional[Callable] = None, **kwargs):
        self.trackers = [ByteTrack(device=device, **kwargs) for _ in range(n_clips)]
        self._device = device
        self._batched_cost = batched_cost        # list of (a, b, scores|None) -> list of cost matrices; default = K4b

    def reset(self):
        for t in self.trackers:
            t.reset()

    def _costs(self, reqs: List[tuple]) -> List[np.ndarray]:
        if self._batched_cost is not None:
            return self._batched_cost(reqs)
        out: List[Optional[np.ndarray]] = [None] * len(reqs)
        live = []
        for i, (a, b, _s) in enumerate(reqs):
            if len(a) == 0 or len(b) == 0:
                out[i] = np.zeros((len(a), len(b)))
            else:
                live.append(i)
        if not live:
            return out
        import torch
        from .runtime import get_context
        ctx = get_context(self._device)
        # every request of one round has the same dtype pattern (the float32-area quirk of numpy's box_iou_batch)
        a0, b0, s0 = reqs[live[0]]
        flags = (1 if np.asarray(a0).dtype == np.float32 else 0) | (2 if np.asarray(b0).dtype == np.float32 else 0)
        A = np.concatenate([np.asarray(reqs[i][0], np.float64).reshape(-1, 4) for i in live])
        B = np.concatenate([np.asarray(reqs[i][1], np.float64).reshape(-1, 4) for i in live])
        na = np.array([len(reqs[i][0]) for i in live]); nb = np.array([len(reqs[i][1]) for i in live])
        a_off = np.concatenate([[0], np.cumsum(na)]).astype(np.int32)
        b_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int32)
        sizes = na * nb
        out_off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        S = np.concatenate([np.asarray(reqs[i][2], np.float64) for i in live]) if s0 is not None else None
        dev = ctx.device
        cost = ctx.iou_cost(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev),
                            torch.from_numpy(S).to(dev) if S is not None else None,
                            torch.from_numpy(a_off).to(dev), torch.from_numpy(b_off).to(dev), torch.from_numpy(out_off).to(dev),
                            len(live), int(na.max()), int(nb.max()), int(sizes.sum()), flags).cpu().numpy()
        for k, i in enumerate(live):
            out[i] = cost[out_off[k]: out_off[k] + sizes[k]].reshape(na[k], nb[k])
        return out

    def update_with_detections(self, detections: Sequence[Detections]) -> List[Detections]:
        """One frame of every clip: detections[i] belongs to clip i.  Returns the tracked detections per clip."""
        assert len(detections) == len(self.trackers)
        gens = [t._update_gen(d) for t, d in zip(self.trackers, detections)]
        results: List[Optional[Detections]] = [None] * len(gens)
        pending = {}
        for i, g in enumerate(gens):
            try:
                pending[i] = next(g)
            except StopIteration as stop:
                results[i] = stop.value
        while pending:
            order = sorted(pending)
            # clips whose request has a different dtype/score pattern than the first one wait for the next batch
            sig = lambda r: (np.asarray(r[0]).dtype == np.float32, np.asarray(r[1]).dtype == np.float32, r[2] is not None)
            first = sig(pending[order[0]])
            batch = [i for i in order if sig(pending[i]) == first]
            costs = self._costs([pending[i] for i in batch])
            for i, c in zip(batch, costs):
                try:
                    pending[i] = gens[i].send(c)
                except StopIteration as stop:
                    results[i] = stop.value
                    del pending[i]
        return results
