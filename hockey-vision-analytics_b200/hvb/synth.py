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


def planted_entries(rng: np.random.Generator, level_hw: Sequence[Tuple[int, int]], nc: int, gt_boxes: np.ndarray,
                    gt_cls: np.ndarray, conf_lo: float = 0.45, conf_hi: float = 0.97, dup: int = 3,
                    background_logit: float = -7.0):
    """The planted anchors of one image as a list of (level, gy, gx, float32[64+nc]) in planting order (a later entry
    on the same anchor replaces an earlier one).

    For every ground-truth box (letterboxed-image xyxy) the level whose stride fits it is chosen, and `dup`
    neighbouring anchors get DFL logits peaked so that the decoded box is (a jittered copy of) the target and a class
    logit giving a DISTINCT confidence in (conf_lo, conf_hi) — duplicates exercise NMS suppression."""
    strides = (8, 16, 32)
    entries = []
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
            vec = np.empty(64 + nc, np.float32)
            for side in range(4):
                k0 = int(np.floor(d[side])); f = d[side] - k0
                logits = np.full(16, -12.0, np.float32)
                logits[k0] = np.log(max(1 - f, 1e-6)) + 6.0
                logits[min(k0 + 1, 15)] = max(logits[min(k0 + 1, 15)], np.log(max(f, 1e-6)) + 6.0)
                vec[side * 16:(side + 1) * 16] = logits
            p = confs[ci]; ci += 1
            vec[64:] = background_logit
            vec[64 + int(c)] = np.log(p / (1 - p))
            entries.append((lvl, gy, gx, vec))
    return entries


def planted_head(rng: np.random.Generator, level_hw: Sequence[Tuple[int, int]], nc: int, gt_boxes: np.ndarray,
                 gt_cls: np.ndarray, conf_lo: float = 0.45, conf_hi: float = 0.97, dup: int = 3,
                 background_logit: float = -7.0) -> List[np.ndarray]:
    """Raw Detect head tensors float32[64+nc, H_i, W_i] (one image) with planted boxes (planted_entries); everything
    else is background (sigmoid ~ 1e-3) with random DFL."""
    out = []
    for (hh, ww) in level_hw:
        t = np.empty((64 + nc, hh, ww), np.float32)
        t[:64] = rng.normal(0, 1, size=(64, hh, ww))
        t[64:] = rng.normal(background_logit, 0.3, size=(nc, hh, ww))
        out.append(t)
    for lvl, gy, gx, vec in planted_entries(rng, level_hw, nc, gt_boxes, gt_cls, conf_lo, conf_hi, dup, background_logit):
        out[lvl][:, gy, gx] = vec
    return out


# ------------------------------------------------------------------ planted detections on top of a real forward
def letterbox_geometry(h: int, w: int, imgsz: int, auto: bool = True, stride: int = 32):
    """(out_h, out_w, gain, pad_x, pad_y) of ultralytics' LetterBox for an h x w image — only used to PLACE synthetic
    boxes on the letterboxed grid (the letterboxing itself is K1's job)."""
    r = min(imgsz / h, imgsz / w)
    uw, uh = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - uw, imgsz - uh
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return uh + top + bottom, uw + left + right, min((uh + top + bottom) / h, (uw + left + right) / w), left, top


def slice_offsets(w: int, h: int, slice_wh=(640, 640), overlap_wh=(128, 128)) -> np.ndarray:
    """Tile rectangles int[n,4] of the slicer (stride = slice - overlap, row-major over y then x, edge tiles clipped)."""
    xs = np.arange(0, w, slice_wh[0] - overlap_wh[0])
    ys = np.arange(0, h, slice_wh[1] - overlap_wh[1])
    x0, y0 = np.meshgrid(xs, ys)
    return np.stack([x0, y0, np.clip(x0 + slice_wh[0], 0, w), np.clip(y0 + slice_wh[1], 0, h)], -1).reshape(-1, 4)


class PlantedOverlay:
    """Sparse planted Detect-head values for a sequence of frames (SURVEY.md §8d "planted mode").

    Random-init YOLO emits nothing above conf 0.4 (H6), so benchmarks and tests that want the post-processing stages to
    do real work overwrite a few anchors of the raw head tensors — AFTER the real forward — with planted_entries: the
    same (frame, tile, level, gy, gx) -> float32[64+nc] table is applied by the CPU reference arm (`apply_host`, on
    the heads of one image) and by the GPU arm (`DeviceOverlay`, a handful of index_put launches on the device heads),
    so both arms decode identical planted candidates on top of their own forward's background.

    tile = 0 for whole-frame detection; for the sliced path a box is planted in every tile that contains at least
    `min_visible` of its area (boxes in the overlap of neighbouring tiles become cross-slice duplicates)."""

    def __init__(self, nc: int):
        self.nc = nc
        self.entries = {}                       # (frame, tile) -> {(lvl, gy, gx): vec}
        self.level_hw = {}                      # (frame, tile) -> [(h, w)] * 3

    def plant(self, frame: int, tile: int, level_hw, entries) -> None:
        d = self.entries.setdefault((frame, tile), {})
        self.level_hw[(frame, tile)] = [tuple(x) for x in level_hw]
        for lvl, gy, gx, vec in entries:
            d[(lvl, gy, gx)] = vec              # last one wins, like sequential writes into a dense tensor

    @classmethod
    def whole_frame(cls, seed: int, frame_hw, imgsz: int, boxes_per_frame, cls_per_frame, nc: int, **kw) -> "PlantedOverlay":
        h, w = frame_hw
        oh, ow, gain, px, py = letterbox_geometry(h, w, imgsz)
        lv = [(oh // s, ow // s) for s in (8, 16, 32)]
        ov, rng = cls(nc), np.random.default_rng(seed)
        for f, (b, c) in enumerate(zip(boxes_per_frame, cls_per_frame)):
            gt = np.asarray(b, np.float64).reshape(-1, 4) * gain + np.array([px, py, px, py])
            ov.plant(f, 0, lv, planted_entries(rng, lv, nc, gt, c, **kw))
        return ov

    @classmethod
    def sliced(cls, seed: int, frame_hw, tile_imgsz: int, boxes_per_frame, cls_per_frame, nc: int, slice_wh=(640, 640),
               overlap_wh=(128, 128), min_visible: float = 0.6, **kw) -> "PlantedOverlay":
        h, w = frame_hw
        offs = slice_offsets(w, h, slice_wh, overlap_wh)
        ov, rng = cls(nc), np.random.default_rng(seed)
        for f, (b, c) in enumerate(zip(boxes_per_frame, cls_per_frame)):
            b = np.asarray(b, np.float64).reshape(-1, 4)
            c = np.asarray(c)
            for t, (x0, y0, x1, y1) in enumerate(offs):
                iw = np.clip(np.minimum(b[:, 2], x1) - np.maximum(b[:, 0], x0), 0, None)
                ih = np.clip(np.minimum(b[:, 3], y1) - np.maximum(b[:, 1], y0), 0, None)
                area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
                sel = iw * ih >= min_visible * area
                if not sel.any():
                    continue
                oh, ow, gain, px, py = letterbox_geometry(int(y1 - y0), int(x1 - x0), tile_imgsz)
                lv = [(oh // s, ow // s) for s in (8, 16, 32)]
                loc = np.clip(b[sel], [x0, y0, x0, y0], [x1, y1, x1, y1]) - np.array([x0, y0, x0, y0])
                ov.plant(f, t, lv, planted_entries(rng, lv, nc, loc * gain + np.array([px, py, px, py]), c[sel], **kw))
        return ov

    def count(self, frame: int) -> int:
        return sum(len(d) for (f, _t), d in self.entries.items() if f == frame)

    def apply_host(self, heads, frame: int, tile: int = 0, batch_index: int = 0) -> None:
        """In place on the 3 raw head tensors [B, 64+nc, H_i, W_i] (numpy arrays or CPU torch tensors) of one image."""
        for (lvl, gy, gx), vec in self.entries.get((frame, tile), {}).items():
            t = heads[lvl]
            if isinstance(t, np.ndarray):
                t[batch_index, :, gy, gx] = vec
            else:
                import torch
                t[batch_index, :, gy, gx] = torch.from_numpy(vec)

    def to_device(self, device, schedule, plan=None) -> "DeviceOverlay":
        return DeviceOverlay(self, device, schedule, plan)


class DeviceOverlay:
    """PlantedOverlay on the device, as a `Detector.head_hook`.

    `schedule` is the cyclic list of chunks the detector is going to see: schedule[k] = frame ids of the k-th call.
    `begin_chunk()` (outside any CUDA graph) copies chunk k's index / value tables into FIXED device buffers with
    device-to-device copies on the current stream; `__call__(heads, cls_index)` (capturable: contents-only dependence)
    scatters them into the heads with index_put.  Shorter tables are padded with repeats of their first entry (same
    value to the same anchor); a table with no entry at all writes the head's current values back."""

    def __init__(self, overlay: PlantedOverlay, device, schedule, plan=None):
        import torch
        self.nc, self.device = overlay.nc, torch.device(device)
        self.schedule = [list(c) for c in schedule]
        self.k = -1
        # (shape class, level) -> per chunk list of (batch_index, gy, gx, vec)
        tiles_of = None
        if plan is not None:
            tpf = plan.tiles_per_frame
            tiles_of = {}
            for t in plan.tiles:
                pos = int(t["frame"])
                tiles_of.setdefault(pos, []).append((int(t["tile"]), int(t["cls"]), int(t["batch_index"])))
        rows = {}
        for k, chunk in enumerate(self.schedule):
            for pos, f in enumerate(chunk):
                places = tiles_of[pos] if tiles_of is not None else [(0, 0, pos)]
                for tile, c, bi in places:
                    for (lvl, gy, gx), vec in overlay.entries.get((f, tile), {}).items():
                        rows.setdefault((c, lvl), {}).setdefault(k, []).append((bi, gy, gx, vec))
        self.tab, self.buf = {}, {}
        nk = len(self.schedule)
        for key, per_chunk in rows.items():
            cap = max(len(v) for v in per_chunk.values())
            idx = np.zeros((nk, 3, cap), np.int64)
            val = np.zeros((nk, cap, 64 + self.nc), np.float32)
            pad = np.ones((nk, 1), bool)
            for k, v in per_chunk.items():
                v = v + [v[0]] * (cap - len(v))
                idx[k] = np.array([[e[0] for e in v], [e[1] for e in v], [e[2] for e in v]])
                val[k] = np.stack([e[3] for e in v])
                pad[k] = False
            self.tab[key] = tuple(torch.from_numpy(a).to(self.device) for a in (idx, val, pad))
            self.buf[key] = (torch.zeros((3, cap), dtype=torch.int64, device=self.device),
                             torch.zeros((cap, 64 + self.nc), dtype=torch.float32, device=self.device),
                             torch.ones((1,), dtype=torch.bool, device=self.device))

    def begin_chunk(self, n_frames: int, repeat: bool = False) -> None:
        """repeat=True: the same chunk again (an overflow retry, a graph warm-up)."""
        if not (repeat and self.k >= 0):
            self.k = (self.k + 1) % len(self.schedule)
        if len(self.schedule[self.k]) != n_frames:
            raise ValueError("overlay schedule expects a chunk of %d frames, got %d" % (len(self.schedule[self.k]), n_frames))
        for key, (idx, val, pad) in self.tab.items():
            b = self.buf[key]
            b[0].copy_(idx[self.k]); b[1].copy_(val[self.k]); b[2].copy_(pad[self.k])

    def __call__(self, heads, cls_index: int = 0):
        import torch
        split = hasattr(heads, "box")
        for lvl in range(3):
            key = (cls_index, lvl)
            if key not in self.buf:
                continue
            idx, val, pad = self.buf[key]
            b, y, x = idx[0], idx[1], idx[2]
            if split:
                for t, v in ((heads.box[lvl], val[:, :64]), (heads.cls[lvl], val[:, 64:])):
                    t[b, :, y, x] = torch.where(pad, t[b, :, y, x], v)
            else:
                t = heads[lvl]
                t[b, :, y, x] = torch.where(pad, t[b, :, y, x], val)
        return heads


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
