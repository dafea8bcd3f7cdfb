"""Seeded synthetic inputs for parity tests and benchmarks: rink-like frames with
coloured player rectangles and a puck, planted-box YOLOv8 head tensors, random boxes.  There are
no videos or checkpoints offline, so every number in this repo is measured on these."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

TEAM_BGR = ((235, 235, 235), (40, 40, 200))   # team 0 white-ish, team 1 red-ish (BGR)


def rink_frame(rng: np.random.Generator, h: int = 1080, w: int = 1920, n_players: int = 12, scale: float = 1.0,
               centres=None):
    """One BGR uint8 frame + player boxes float32[n,4] + team labels int[n] + puck box float32[4]."""
    base = rng.normal(235.0, 8.0, size=(h, w, 1)).astype(np.float32)
    frame = np.clip(base + rng.normal(0, 2.0, size=(h, w, 3)).astype(np.float32), 0, 255).astype(np.uint8)
    boxes, teams = [], []
    for i in range(n_players):
        ph = int(rng.integers(100, 251) * scale)
        pw = int(rng.integers(40, 111) * scale)
        if centres is None:
            cx = int(rng.integers(pw // 2 + 1, w - pw // 2 - 1))
            cy = int(rng.integers(ph // 2 + 1, h - ph // 2 - 1))
        else:
            cx = int(np.clip(centres[i][0], pw // 2 + 1, w - pw // 2 - 2))
            cy = int(np.clip(centres[i][1], ph // 2 + 1, h - ph // 2 - 2))
        x0, y0 = cx - pw // 2, cy - ph // 2
        team = i % 2
        col = np.array(TEAM_BGR[team], np.float32)
        patch = np.clip(col + rng.normal(0, 25.0, size=(ph, pw, 3)).astype(np.float32), 0, 255).astype(np.uint8)
        frame[y0:y0 + ph, x0:x0 + pw] = patch
        boxes.append([x0, y0, x0 + pw, y0 + ph])
        teams.append(team)
    pr = int(rng.integers(3, 6) * scale)
    px, py = int(rng.integers(pr + 1, w - pr - 1)), int(rng.integers(pr + 1, h - pr - 1))
    yy, xx = np.ogrid[-pr:pr + 1, -pr:pr + 1]
    disc = (xx * xx + yy * yy) <= pr * pr
    sub = frame[py - pr:py + pr + 1, px - pr:px + pr + 1]
    sub[disc] = (20, 20, 20)
    puck = np.array([px - pr, py - pr, px + pr + 1, py + pr + 1], np.float32)
    return frame, np.asarray(boxes, np.float32).reshape(-1, 4), np.asarray(teams, int), puck


def rink_clip(seed: int, n_frames: int, h: int = 1080, w: int = 1920, n_players: int = 12, scale: float = 1.0):
    """A clip whose players drift <= 8 px / frame so a tracker keeps identities."""
    rng = np.random.default_rng(seed)
    centres = np.stack([rng.integers(150, w - 150, n_players), rng.integers(200, h - 200, n_players)], 1).astype(np.float64)
    frames, boxes, teams, pucks = [], [], [], []
    for _ in range(n_frames):
        centres += rng.uniform(-8, 8, size=centres.shape)
        frng = np.random.default_rng(rng.integers(1 << 31))
        f, b, t, p = rink_frame(frng, h, w, n_players, scale, centres)
        frames.append(f); boxes.append(b); teams.append(t); pucks.append(p)
    return np.stack(frames), boxes, teams, pucks


def random_frames(seed: int, n: int, h: int, w: int) -> np.ndarray:
    """Uniform-random frames: worst case for histogram contention and resize parity."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def random_boxes(rng: np.random.Generator, n: int, w: float = 1920, h: float = 1080, smin: float = 10, smax: float = 70,
                 dtype=np.float32):
    """n xyxy boxes with uniform centres and sizes, plus DISTINCT scores in (0.05, 0.99)."""
    cx, cy = rng.uniform(0, w, n), rng.uniform(0, h, n)
    bw, bh = rng.uniform(smin, smax, n), rng.uniform(smin, smax, n)
    boxes = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).astype(dtype)
    scores = rng.permutation(np.linspace(0.05, 0.99, n)).astype(np.float32)
    return boxes, scores


def planted_head(rng: np.random.Generator, level_hw: Sequence[Tuple[int, int]], nc: int, gt_boxes: np.ndarray,
                 gt_cls: np.ndarray, conf_lo: float = 0.45, conf_hi: float = 0.97, dup: int = 3,
                 background_logit: float = -7.0) -> List[np.ndarray]:
    """Raw Detect head tensors float32[64+nc, H_i, W_i] (one image) with planted boxes.

    For every ground-truth box (letterboxed-image xyxy) the level whose stride fits it is chosen,
    and `dup` neighbouring anchors get DFL logits peaked so that the decoded box is (a jittered copy
    of) the target and a class logit giving a DISTINCT confidence in (conf_lo, conf_hi) — duplicates
    exercise NMS suppression.  Everything else is background (sigmoid ~ 1e-3) with random DFL."""
    strides = (8, 16, 32)
    out = []
    for (hh, ww) in level_hw:
        t = np.empty((64 + nc, hh, ww), np.float32)
        t[:64] = rng.normal(0, 1, size=(64, hh, ww))
        t[64:] = rng.normal(background_logit, 0.3, size=(nc, hh, ww))
        out.append(t)
    n_plant = len(gt_boxes) * dup
    confs = rng.permutation(np.linspace(conf_lo, conf_hi, max(n_plant, 1)))
    ci = 0
    for box, c in zip(gt_boxes, gt_cls):
        x1, y1, x2, y2 = [float(v) for v in box]
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        lvl = 2
        for li, s in enumerate(strides):
            if max(x2 - x1, y2 - y1) / 2 / s < 13.0:
                lvl = li
                break
        s = strides[lvl]
        hh, ww = level_hw[lvl]
        gx0, gy0 = int(cx / s), int(cy / s)
        for k in range(dup):
            gx = min(max(gx0 + (k % 2), 0), ww - 1)
            gy = min(max(gy0 + (k // 2), 0), hh - 1)
            ax, ay = gx + 0.5, gy + 0.5
            jit = rng.uniform(-0.6, 0.6, 4) if k else np.zeros(4)
            d = np.array([ax - x1 / s, ay - y1 / s, x2 / s - ax, y2 / s - ay]) + jit / s * 4
            d = np.clip(d, 0.05, 14.9)
            for side in range(4):
                k0 = int(np.floor(d[side])); f = d[side] - k0
                logits = np.full(16, -12.0, np.float32)
                logits[k0] = np.log(max(1 - f, 1e-6)) + 6.0
                logits[min(k0 + 1, 15)] = max(logits[min(k0 + 1, 15)], np.log(max(f, 1e-6)) + 6.0)
                out[lvl][side * 16:(side + 1) * 16, gy, gx] = logits
            p = confs[ci]; ci += 1
            out[lvl][64:, gy, gx] = background_logit
            out[lvl][64 + int(c), gy, gx] = np.log(p / (1 - p))
    return out


def pack_crops(crops: Sequence[np.ndarray]):
    """List of (possibly non-contiguous) uint8[h,w,3] crops -> (packed bytes uint8[N], CROP_DESC-compatible
    rows (offset, pitch, h, w)).  One memcpy per crop row block; this is the host-side 'pack once'
    step for the reference's list-of-views call surface."""
    total = sum(int(c.shape[0]) * int(c.shape[1]) * 3 for c in crops)
    buf = np.empty((max(total, 1),), np.uint8)
    desc = np.zeros((len(crops), 4), np.int64)
    off = 0
    for i, c in enumerate(crops):
        h, w = int(c.shape[0]), int(c.shape[1])
        n = h * w * 3
        if n:
            buf[off:off + n].reshape(h, w, 3)[...] = c
        desc[i] = (off, w * 3, h, w)
        off += n
    return buf, desc
