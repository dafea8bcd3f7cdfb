"""Bit-exact numpy restatements of the OpenCV / Pillow integer arithmetic the
reference's hot path reaches (TEST INFRASTRUCTURE — see oracle/__init__.py).

Reference call sites:
  cv2.cvtColor(BGR2HSV / BGR2LAB)   hockey/common/team_hybrid.py:97-98
  cv2.calcHist                      hockey/common/team_hybrid.py:101-103
  transforms.Resize((128, 64))      hockey/common/team_hybrid.py:31-36  (Pillow BILINEAR, antialiased)
  cv2.resize(INTER_LINEAR)          inside ultralytics LetterBox, reached from hockey/main.py:179-184

The algorithms live in third-party libraries that are not vendored by the
reference (opencv-python, Pillow; versions unpinned there; 4.13.0 / 12.2.0 in
this image).  Each restatement below is pinned by tests/test_oracle_cv_exact.py
against the installed library (full 2^24 colour cube for the colour
conversions).
"""
from __future__ import annotations

import functools
import numpy as np

# ----------------------------------------------------------------------------------------------
# A1. BGR -> HSV (uint8, H in [0,180))            OpenCV color_hsv: RGB2HSV_b
# ----------------------------------------------------------------------------------------------
HSV_SHIFT = 12


@functools.lru_cache(maxsize=None)
def hsv_tables():
    """(sdiv, hdiv) int32[256]:  sdiv[i]=rint((255<<12)/i), hdiv[i]=rint((180<<12)/(6 i)), [0]=0."""
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, np.int32)
    hdiv = np.zeros(256, np.int32)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int32)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return sdiv, hdiv


def bgr2hsv(bgr: np.ndarray) -> np.ndarray:
    """uint8[...,3] BGR -> uint8[...,3] HSV, identical to cv2.cvtColor(COLOR_BGR2HSV)."""
    sdiv, hdiv = hsv_tables()
    x = np.asarray(bgr)
    b = x[..., 0].astype(np.int32)
    g = x[..., 1].astype(np.int32)
    r = x[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    s = (diff * sdiv[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT  # arithmetic shift on signed
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# A2. BGR -> LAB (uint8)                           OpenCV color_lab: RGB2Lab_b
# ----------------------------------------------------------------------------------------------
LAB_SHIFT = 12
LAB_SHIFT2 = 15
LAB_C = np.array([[1777, 1541, 778], [871, 2929, 296], [73, 448, 3575]], np.int64)  # rows X,Y,Z on (R,G,B)


@functools.lru_cache(maxsize=None)
def lab_tables():
    """(gtab uint16[256], ctab uint16[3072]) built with float32 arithmetic like OpenCV's softfloat."""
    f32 = np.float32
    i = np.arange(256, dtype=np.float32)
    x = i / f32(255.0)
    with np.errstate(all="ignore"):
        hi = np.power((x + f32(0.055)) / f32(1.055), f32(2.4), dtype=np.float32)
    y = np.where(x <= f32(0.04045), x / f32(12.92), hi).astype(np.float32)
    gtab = np.rint(f32(2040.0) * y).astype(np.uint16)

    j = np.arange(3072, dtype=np.float32)
    xs = (j * (f32(1.0) / (f32(255.0) * f32(8.0)))).astype(np.float32)
    lin = (xs * (f32(841.0) / f32(108.0))).astype(np.float32) + (f32(16.0) / f32(116.0))
    cb = np.cbrt(xs).astype(np.float32)
    yc = np.where(xs < f32(216.0) / f32(24389.0), lin.astype(np.float32), cb)
    ctab = np.rint(f32(32768.0) * yc.astype(np.float32)).astype(np.uint16)
    return gtab, ctab


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def bgr2lab(bgr: np.ndarray) -> np.ndarray:
    """uint8[...,3] BGR -> uint8[...,3] LAB, identical to cv2.cvtColor(COLOR_BGR2LAB)."""
    gtab, ctab = lab_tables()
    x = np.asarray(bgr)
    B = gtab[x[..., 0]].astype(np.int64)
    G = gtab[x[..., 1]].astype(np.int64)
    R = gtab[x[..., 2]].astype(np.int64)
    C = LAB_C
    fX = ctab[_descale(R * C[0, 0] + G * C[0, 1] + B * C[0, 2], LAB_SHIFT)].astype(np.int64)
    fY = ctab[_descale(R * C[1, 0] + G * C[1, 1] + B * C[1, 2], LAB_SHIFT)].astype(np.int64)
    fZ = ctab[_descale(R * C[2, 0] + G * C[2, 1] + B * C[2, 2], LAB_SHIFT)].astype(np.int64)
    L = _descale(296 * fY - 1336934, LAB_SHIFT2)
    a = _descale(500 * (fX - fY) + 128 * 32768, LAB_SHIFT2)
    b = _descale(200 * (fY - fZ) + 128 * 32768, LAB_SHIFT2)
    return np.clip(np.stack([L, a, b], axis=-1), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# A3. calcHist on one uint8 channel with a uniform range [0, hi)
# ----------------------------------------------------------------------------------------------
def calc_hist_u8(channel: np.ndarray, bins: int, hi: int) -> np.ndarray:
    """float32[bins] counts like cv2.calcHist([img],[c],None,[bins],[0,hi]).flatten().

    OpenCV maps v -> floor(v * bins / hi) for v < hi and drops the rest; for the three
    specs the reference uses this is v//10 (18 bins / 180) and v>>5 (8 bins / 256).
    """
    v = np.asarray(channel).astype(np.int64).ravel()
    v = v[v < hi]
    idx = (v * bins) // hi
    return np.bincount(idx, minlength=bins).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# A4. Pillow Image.resize(BILINEAR) on uint8 (what transforms.Resize does on a PIL image)
# ----------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_coeffs(in_size: int, out_size: int):
    """Per-output (xmin, n, int32 coeffs[ksize]) of Pillow's precompute_coeffs + normalize_coeffs_8bpc
    for the bilinear (triangle, support 1.0) filter over the box [0, in_size)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + xmin - center + 0.5) * ss
            if t < 0.0:
                t = -t
            wx = 1.0 - t if t < 1.0 else 0.0
            w[x] = wx
            ww += wx
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            p = w[x] * (1 << PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + p) if w[x] < 0 else int(0.5 + p)
        bounds[xx] = (xmin, n)
    return bounds, kk


def _pil_pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) of a uint8[H,W,C] image."""
    in_size = img.shape[axis]
    bounds, kk = pil_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)           # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        xmin, n = bounds[o]
        acc = np.tensordot(kk[o, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0))
        acc = (acc + (1 << (PIL_PRECISION_BITS - 1))) >> PIL_PRECISION_BITS
        out[o] = np.clip(acc, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8[H,W,C] -> uint8[out_h,out_w,C], identical to PIL.Image.fromarray(img).resize((out_w,out_h), BILINEAR).

    Two separable passes with a uint8 intermediate.  Pass order: horizontal then vertical,
    unless vertical_first(...) says otherwise (see there).  A pass whose size is unchanged is skipped.
    """
    img = np.ascontiguousarray(img)
    in_h, in_w = img.shape[:2]
    need_h = out_w != in_w
    need_v = out_h != in_h
    if pil_vertical_first(in_w, in_h, out_w, out_h):
        if need_v:
            img = _pil_pass(img, out_h, 0)
        if need_h:
            img = _pil_pass(img, out_w, 1)
    else:
        if need_h:
            img = _pil_pass(img, out_w, 1)
        if need_v:
            img = _pil_pass(img, out_h, 0)
    return img


def pil_vertical_first(in_w: int, in_h: int, out_w: int, out_h: int) -> bool:
    """Pass-order rule of the installed Pillow (12.2.0), found empirically and
    re-checked by tests/test_oracle_cv_exact.py: horizontal first except when in_h > 100 * in_w."""
    return in_h > 100 * in_w


def mnv3_preprocess(roi_bgr: np.ndarray) -> np.ndarray:
    """team_hybrid.py:31-36 on one jersey ROI: ToPILImage -> Resize((128,64)) -> ToTensor -> Normalize.
    The BGR crop is fed to the RGB statistics unswapped (reference quirk).
    Returns float32[3,128,64]."""
    u8 = pil_resize_bilinear(roi_bgr, 64, 128)
    x = u8.astype(np.float32) / np.float32(255.0)
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    std = np.array([0.229, 0.224, 0.225], np.float32)
    x = (x - mean) / std
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# A5. cv2.resize(INTER_LINEAR) on uint8 (LetterBox path)
# ----------------------------------------------------------------------------------------------
CV_COEF_BITS = 11
CV_COEF_SCALE = 1 << CV_COEF_BITS


def cv_linear_coeffs(src: int, dst: int, clamp_fraction: bool):
    """(idx int32[dst], c0 int32[dst], c1 int32[dst]).  `clamp_fraction` = horizontal-axis behaviour
    (fraction zeroed at the borders); the vertical axis keeps the fraction and clamps only indices."""
    scale = float(src) / dst
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo = s < 0
        hi = s >= src - 1
        f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        s = np.where(lo, 0, np.where(hi, src - 1, s)).astype(np.int32)
    c0 = np.rint((np.float32(1.0) - f) * np.float32(CV_COEF_SCALE)).astype(np.int32)
    c1 = np.rint(f * np.float32(CV_COEF_SCALE)).astype(np.int32)
    return s, c0, c1


def cv_resize_linear(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """uint8[H,W,C] -> uint8[dst_h,dst_w,C], identical to cv2.resize(img,(dst_w,dst_h),interpolation=INTER_LINEAR)."""
    img = np.ascontiguousarray(img)
    src_h, src_w = img.shape[:2]
    if src_h == dst_h and src_w == dst_w:
        return img.copy()
    if src_w == 2 * dst_w and src_h == 2 * dst_h:
        # OpenCV switches INTER_LINEAR to 2x2 area averaging when both factors are exactly 2.
        p = img.astype(np.int32)
        out = (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8)
    sx, a0, a1 = cv_linear_coeffs(src_w, dst_w, clamp_fraction=True)
    sy, b0, b1 = cv_linear_coeffs(src_h, dst_h, clamp_fraction=False)
    sx1 = np.minimum(sx + 1, src_w - 1)
    S = img.astype(np.int32)
    rows = S[:, sx, :] * a0[None, :, None] + S[:, sx1, :] * a1[None, :, None]   # [src_h, dst_w, C] int32
    r0 = np.clip(sy, 0, src_h - 1)
    r1 = np.clip(sy + 1, 0, src_h - 1)
    R0 = rows[r0] >> 4
    R1 = rows[r1] >> 4
    out = (((b0[:, None, None] * R0) >> 16) + ((b1[:, None, None] * R1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
