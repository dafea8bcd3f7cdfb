"""CPU restatement of supervision's ByteTrack (tracker/byte_tracker/{core,matching,kalman_filter}.py),
the tracker the reference constructs at hockey/main.py:162-168 and 207-211 and steps at :228, :265.
TEST INFRASTRUCTURE — see oracle/__init__.py.  PARITY UNPINNED (supervision is absent from
/root/reference and from this image); restated object-by-object from the published algorithm
(SURVEY.md App. B2): xyah constant-velocity Kalman filter (std weights 1/20, 1/160), three
associations (high-score with fuse_score at minimum_matching_threshold, low-score at 0.5,
unconfirmed at 0.7), Hungarian assignment with scipy, lost/removed bookkeeping, external ids
issued after minimum_consecutive_frames.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

from .supervision_restated import box_iou_batch, fuse_score as _fuse, iou_distance as _iou_distance

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


class KalmanFilter:
    def __init__(self):
        ndim, dt = 4, 1.0
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    def initiate(self, measurement):
        mean = np.r_[measurement, np.zeros_like(measurement)]
        h = measurement[3]
        std = [2 * self._std_weight_position * h, 2 * self._std_weight_position * h, 1e-2, 2 * self._std_weight_position * h,
               10 * self._std_weight_velocity * h, 10 * self._std_weight_velocity * h, 1e-5, 10 * self._std_weight_velocity * h]
        return mean, np.diag(np.square(std))

    def project(self, mean, covariance):
        h = mean[3]
        std = [self._std_weight_position * h, self._std_weight_position * h, 1e-1, self._std_weight_position * h]
        innovation_cov = np.diag(np.square(std))
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot((self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + innovation_cov

    def multi_predict(self, mean, covariance):
        std_pos = [self._std_weight_position * mean[:, 3], self._std_weight_position * mean[:, 3],
                   1e-2 * np.ones_like(mean[:, 3]), self._std_weight_position * mean[:, 3]]
        std_vel = [self._std_weight_velocity * mean[:, 3], self._std_weight_velocity * mean[:, 3],
                   1e-5 * np.ones_like(mean[:, 3]), self._std_weight_velocity * mean[:, 3]]
        sqr = np.square(np.r_[std_pos, std_vel]).T
        motion_cov = np.asarray([np.diag(sqr[i]) for i in range(len(mean))])
        mean = np.dot(mean, self._motion_mat.T)
        left = np.dot(self._motion_mat, covariance).transpose((1, 0, 2))
        covariance = np.dot(left, self._motion_mat.T) + motion_cov
        return mean, covariance

    def update(self, mean, covariance, measurement):
        projected_mean, projected_cov = self.project(mean, covariance)
        chol_factor, lower = scipy.linalg.cho_factor(projected_cov, lower=True, check_finite=False)
        kalman_gain = scipy.linalg.cho_solve((chol_factor, lower), np.dot(covariance, self._update_mat.T).T,
                                             check_finite=False).T
        innovation = measurement - projected_mean
        new_mean = mean + np.dot(innovation, kalman_gain.T)
        new_covariance = covariance - np.linalg.multi_dot((kalman_gain, projected_cov, kalman_gain.T))
        return new_mean, new_covariance


class IdCounter:
    NO_ID = -1

    def __init__(self, start_id: int = 0):
        self.start_id = start_id
        self._id = start_id

    def new_id(self):
        i = self._id
        self._id += 1
        return i


class STrack:
    def __init__(self, tlwh, score, minimum_consecutive_frames, shared_kalman, internal_counter, external_counter):
        self.state = NEW
        self.is_activated = False
        self.start_frame = 0
        self.frame_id = 0
        self._tlwh = np.asarray(tlwh, dtype=np.float32)
        self.kalman_filter = None
        self.shared_kalman = shared_kalman
        self.mean, self.covariance = None, None
        self.score = score
        self.tracklet_len = 0
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self.internal_counter, self.external_counter = internal_counter, external_counter
        self.internal_track_id = IdCounter.NO_ID
        self.external_track_id = IdCounter.NO_ID

    @staticmethod
    def multi_predict(stracks, shared_kalman):
        if len(stracks) > 0:
            multi_mean = np.asarray([st.mean.copy() for st in stracks])
            multi_covariance = np.asarray([st.covariance for st in stracks])
            for i, st in enumerate(stracks):
                if st.state != TRACKED:
                    multi_mean[i][7] = 0
            multi_mean, multi_covariance = shared_kalman.multi_predict(multi_mean, multi_covariance)
            for i, (mean, cov) in enumerate(zip(multi_mean, multi_covariance)):
                stracks[i].mean = mean
                stracks[i].covariance = cov

    def activate(self, kalman_filter, frame_id):
        self.kalman_filter = kalman_filter
        self.internal_track_id = self.internal_counter.new_id()
        self.mean, self.covariance = self.kalman_filter.initiate(self.tlwh_to_xyah(self._tlwh))
        self.tracklet_len = 0
        self.state = TRACKED
        if frame_id == 1:
            self.is_activated = True
        if self.minimum_consecutive_frames == 1:
            self.external_track_id = self.external_counter.new_id()
        self.frame_id = frame_id
        self.start_frame = frame_id

    def re_activate(self, new_track, frame_id):
        self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.tracklet_len = 0
        self.state = TRACKED
        self.frame_id = frame_id
        self.score = new_track.score

    def update(self, new_track, frame_id):
        self.frame_id = frame_id
        self.tracklet_len += 1
        self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.state = TRACKED
        if self.tracklet_len == self.minimum_consecutive_frames:
            self.is_activated = True
            if self.external_track_id == IdCounter.NO_ID:
                self.external_track_id = self.external_counter.new_id()
        self.score = new_track.score

    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    @property
    def tlbr(self):
        ret = self.tlwh.copy()
        ret[2:] += ret[:2]
        return ret

    @staticmethod
    def tlwh_to_xyah(tlwh):
        ret = np.asarray(tlwh).copy()
        ret[:2] += ret[2:] / 2
        ret[2] /= ret[3]
        return ret

    @staticmethod
    def tlbr_to_tlwh(tlbr):
        ret = np.asarray(tlbr).copy()
        ret[2:] -= ret[:2]
        return ret


def linear_assignment(cost_matrix: np.ndarray, thresh: float):
    if cost_matrix.size == 0:
        return np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1]))
    cost_matrix = cost_matrix.copy()
    cost_matrix[cost_matrix > thresh] = thresh + 1e-4
    row_ind, col_ind = linear_sum_assignment(cost_matrix)
    indices = np.column_stack((row_ind, col_ind))
    matched_cost = cost_matrix[tuple(zip(*indices))]
    matches = indices[matched_cost <= thresh]
    unmatched_a = tuple(sorted(set(range(cost_matrix.shape[0])) - set(matches[:, 0])))
    unmatched_b = tuple(sorted(set(range(cost_matrix.shape[1])) - set(matches[:, 1])))
    return matches, unmatched_a, unmatched_b


def iou_distance(atracks: List[STrack], btracks: List[STrack]) -> np.ndarray:
    atlbrs = [t.tlbr for t in atracks]
    btlbrs = [t.tlbr for t in btracks]
    if len(atlbrs) == 0 or len(btlbrs) == 0:
        return np.zeros((len(atlbrs), len(btlbrs)), dtype=np.float32)
    return _iou_distance(np.asarray(atlbrs), np.asarray(btlbrs))


def fuse_score(cost_matrix, detections: List[STrack]):
    if cost_matrix.size == 0:
        return cost_matrix
    return _fuse(cost_matrix, np.array([d.score for d in detections]))


def joint_tracks(a, b):
    seen, out = set(), []
    for t in a:
        seen.add(t.internal_track_id)
        out.append(t)
    for t in b:
        if t.internal_track_id not in seen:
            seen.add(t.internal_track_id)
            out.append(t)
    return out


def sub_tracks(a, b):
    ids = {t.internal_track_id for t in b}
    return [t for t in a if t.internal_track_id not in ids]


def remove_duplicate_tracks(a, b):
    pairwise = iou_distance(a, b)
    pairs = np.where(pairwise < 0.15)
    dup_a, dup_b = set(), set()
    for ia, ib in zip(*pairs):
        time_a = a[ia].frame_id - a[ia].start_frame
        time_b = b[ib].frame_id - b[ib].start_frame
        if time_a > time_b:
            dup_b.add(ib)
        else:
            dup_a.add(ia)
    return [t for i, t in enumerate(a) if i not in dup_a], [t for i, t in enumerate(b) if i not in dup_b]


class ByteTrack:
    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: int = 30, minimum_consecutive_frames: int = 1):
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.frame_id = 0
        self.det_thresh = self.track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self.kalman_filter = KalmanFilter()
        self.shared_kalman = KalmanFilter()
        self.tracked_tracks: List[STrack] = []
        self.lost_tracks: List[STrack] = []
        self.removed_tracks: List[STrack] = []
        self.internal_id_counter = IdCounter()
        self.external_id_counter = IdCounter(start_id=1)

    def update_with_detections(self, xyxy: np.ndarray, confidence: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Returns (indices of the kept detections, their tracker ids) — what
        ``detections[detections.tracker_id != -1]`` carries in supervision."""
        tensors = np.hstack((xyxy, confidence[:, np.newaxis]))
        tracks = self.update_with_tensors(tensors)
        if len(tracks) > 0 and len(tensors) > 0:
            det_boxes = np.asarray([d[:4] for d in tensors])
            track_boxes = np.asarray([t.tlbr for t in tracks])
            ious = box_iou_batch(det_boxes, track_boxes)
            matches, _, _ = linear_assignment(1 - ious, 0.5)
            ids = np.full(len(tensors), -1, dtype=int)
            for i_det, i_trk in matches:
                ids[i_det] = int(tracks[i_trk].external_track_id)
            keep = np.nonzero(ids != -1)[0]
            return keep, ids[keep]
        return np.zeros(0, int), np.zeros(0, int)

    def update_with_tensors(self, tensors: np.ndarray) -> List[STrack]:
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        scores, bboxes = tensors[:, 4], tensors[:, :4]
        remain = scores > self.track_activation_threshold
        second = np.logical_and(scores > 0.1, scores < self.track_activation_threshold)
        mk = lambda tlbr, s: STrack(STrack.tlbr_to_tlwh(tlbr), s, self.minimum_consecutive_frames, self.shared_kalman,
                                    self.internal_id_counter, self.external_id_counter)
        detections = [mk(b, s) for b, s in zip(bboxes[remain], scores[remain])]
        unconfirmed = [t for t in self.tracked_tracks if not t.is_activated]
        tracked = [t for t in self.tracked_tracks if t.is_activated]

        pool = joint_tracks(tracked, self.lost_tracks)
        STrack.multi_predict(pool, self.shared_kalman)
        dists = fuse_score(iou_distance(pool, detections), detections)
        matches, u_track, u_detection = linear_assignment(dists, self.minimum_matching_threshold)
        for it, idet in matches:
            trk, det = pool[it], detections[idet]
            if trk.state == TRACKED:
                trk.update(det, self.frame_id)
                activated.append(trk)
            else:
                trk.re_activate(det, self.frame_id)
                refind.append(trk)

        detections_second = [mk(b, s) for b, s in zip(bboxes[second], scores[second])]
        r_tracked = [pool[i] for i in u_track if pool[i].state == TRACKED]
        dists = iou_distance(r_tracked, detections_second)
        matches, u_track2, _ = linear_assignment(dists, 0.5)
        for it, idet in matches:
            trk, det = r_tracked[it], detections_second[idet]
            if trk.state == TRACKED:
                trk.update(det, self.frame_id)
                activated.append(trk)
            else:
                trk.re_activate(det, self.frame_id)
                refind.append(trk)
        for it in u_track2:
            trk = r_tracked[it]
            if trk.state != LOST:
                trk.state = LOST
                lost.append(trk)

        detections = [detections[i] for i in u_detection]
        dists = fuse_score(iou_distance(unconfirmed, detections), detections)
        matches, u_unconfirmed, u_detection = linear_assignment(dists, 0.7)
        for it, idet in matches:
            unconfirmed[it].update(detections[idet], self.frame_id)
            activated.append(unconfirmed[it])
        for it in u_unconfirmed:
            unconfirmed[it].state = REMOVED
            removed.append(unconfirmed[it])

        for inew in u_detection:
            trk = detections[inew]
            if trk.score < self.det_thresh:
                continue
            trk.activate(self.kalman_filter, self.frame_id)
            activated.append(trk)

        for trk in self.lost_tracks:
            if self.frame_id - trk.frame_id > self.max_time_lost:
                trk.state = REMOVED
                removed.append(trk)

        self.tracked_tracks = [t for t in self.tracked_tracks if t.state == TRACKED]
        self.tracked_tracks = joint_tracks(self.tracked_tracks, activated)
        self.tracked_tracks = joint_tracks(self.tracked_tracks, refind)
        self.lost_tracks = sub_tracks(self.lost_tracks, self.tracked_tracks)
        self.lost_tracks.extend(lost)
        self.lost_tracks = sub_tracks(self.lost_tracks, self.removed_tracks)
        self.removed_tracks = removed
        self.tracked_tracks, self.lost_tracks = remove_duplicate_tracks(self.tracked_tracks, self.lost_tracks)
        return [t for t in self.tracked_tracks if t.is_activated]
