"""CPU proofs of the integer / float identities the K1 column loop relies on (csrc/k1_letterbox.cu, round 2):

* the vertical pass of OpenCV's 8-bit bilinear resize, (((c0*h0)>>16) + ((c1*h1)>>16) + 2) >> 2, evaluated as
  byte 1 of  64 * (hi16(c0*h0 + (2<<16)) + hi16(c1*h1))  (two IMAD, one PRMT, one IDP.2A by (64, 64));
* byte -> float32 / 255 as  RZ((2^23 + v) * k - 2^23 * k)  with k = 0x1.010102p-8 (one PRMT + one FFMA.RZ), which must
  equal numpy's float32(v) / float32(255) for every byte.

The GPU tests (tests/test_gpu_letterbox.py) check the kernel against cv2 bit for bit; these checks pin the arithmetic
itself, exhaustively where the domain is small, and run without a GPU.
"""
from fractions import Fraction

import numpy as np


def _vertical_reference(c0, h0, c1, h1):
    return (((c0 * h0) >> 16) + ((c1 * h1) >> 16) + 2) >> 2


def _vertical_kernel(c0, h0, c1, h1):
    p0 = (c0 * h0 + 0x20000) & 0xFFFFFFFF                 # IMAD with the rounding constant riding on the product
    p1 = (c1 * h1) & 0xFFFFFFFF
    hi2_lo, hi2_hi = p0 >> 16, p1 >> 16                  # PRMT 0x7632: the two upper halves in one register
    acc = (64 * hi2_lo + 64 * hi2_hi + 0x4B000000) & 0xFFFFFFFF      # IDP.2A.LO.U16.U8 by bytes (64, 64), accumulator 0x4B000000
    assert np.all((acc >> 24) == 0x4B) and np.all(((acc >> 16) & 0xFF) == 0)   # exponent byte planted, byte 2 clear
    return (acc >> 8) & 0xFF                             # the pixel value is byte 1


def test_vertical_pass_identity_all_coefficients():
    # h = (horizontal dp2a result) >> 4 <= (255 * 2048) >> 4 = 32640; c0 + c1 = 2048 (11-bit fixed point)
    rng = np.random.default_rng(0)
    hs = np.unique(np.concatenate([np.arange(0, 64), np.arange(32640 - 64, 32641), rng.integers(0, 32641, 150)])).astype(np.int64)
    h0, h1 = np.meshgrid(hs, hs, indexing="ij")
    for c0 in range(0, 2049):
        c1 = 2048 - c0
        ref = _vertical_reference(c0, h0, c1, h1)
        got = _vertical_kernel(np.int64(c0), h0, np.int64(c1), h1)
        assert ref.max() <= 255
        assert np.array_equal(ref, got), c0


def test_vertical_pass_bounds():
    # the claims in the kernel comment: products < 2^27 (no carry into the added 2<<16), 64 * sum < 2^16
    assert 2048 * 32640 < 2 ** 27
    s_max = ((2048 * 32640) >> 16) + 2
    assert s_max == 1022 and 64 * s_max < 2 ** 16 and (s_max >> 2) == 255


def _rz_float32(x: Fraction) -> np.float32:
    """Round a non-negative exact rational toward zero to float32."""
    if x == 0:
        return np.float32(0.0)
    e = 0
    while x >= 2:
        x /= 2; e += 1
    while x < 1:
        x *= 2; e -= 1
    mant = int(x * (1 << 23))                             # floor: round toward zero
    return np.float32(np.ldexp(np.float64(mant), e - 23))


def test_byte_over_255_through_fma_rz():
    k = Fraction(float(np.float32(float.fromhex("0x1.010102p-8"))))
    c = Fraction(float(np.float32(float.fromhex("0x1.010102p+15"))))
    assert c == k * (1 << 23)                             # the addend is exactly 2^23 * k, hence representable
    for v in range(256):
        magic = np.array([0x4B000000 | v], dtype=np.uint32).view(np.float32)[0]
        assert float(magic) == float(2 ** 23 + v)          # PRMT drops the byte into the mantissa of 2^23
        exact = Fraction(float(magic)) * k - c             # what the FMA evaluates before its single rounding
        assert exact == v * k
        got = _rz_float32(exact)
        want = np.float32(v) / np.float32(255.0)           # numpy / torch `x / 255` in float32 (round to nearest)
        assert got == want and np.signbit(got) == np.signbit(want), v
