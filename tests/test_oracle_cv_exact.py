"""Pin oracle/cv_exact.py against the installed OpenCV / Pillow."""
import hashlib

import cv2
import numpy as np
import pytest
from PIL import Image

from oracle import cv_exact as cx

HSV_CUBE_SHA = "cc4c8f3a2064ffaed3776170c4dfa02c90011b02dbc7ece07b7fced54069ad55"
LAB_CUBE_SHA = "6777b2103b2347e79cfcaf30f90002ede0142304b76abfd6100310e5ea68480c"


def colour_cube():
    b, g, r = np.meshgrid(*[np.arange(256, dtype=np.uint8)] * 3, indexing="ij")
    return np.stack([b, g, r], -1).reshape(4096, 4096, 3)


def test_table_hashes():
    g, c = cx.lab_tables()
    s, h = cx.hsv_tables()
    assert hashlib.sha256(g.tobytes()).hexdigest() == "8bfeace00785402e67e5c7d4c53961990e4987ccd31692c42d4b080aeb6dc153"
    assert hashlib.sha256(c.tobytes()).hexdigest() == "bda905efdc57563cfecc16da002a0ef811f6ce1a30fe8b20f44f64efbc264e66"
    assert hashlib.sha256(s.tobytes()).hexdigest() == "d013bc0461c36fe6bfc497b492edc49416f380c16bb17c1cd28f245bde45113d"
    assert hashlib.sha256(h.tobytes()).hexdigest() == "701179918a8d0c3aff73d8d60b527718574f9a36dd3985ca0fe60b4fee45ee8d"


def test_hsv_lab_full_cube_matches_cv2():
    cube = colour_cube()
    hsv = cv2.cvtColor(cube, cv2.COLOR_BGR2HSV)
    lab = cv2.cvtColor(cube, cv2.COLOR_BGR2LAB)
    assert hashlib.sha256(hsv.tobytes()).hexdigest() == HSV_CUBE_SHA
    assert hashlib.sha256(lab.tobytes()).hexdigest() == LAB_CUBE_SHA
    # restatement on a strided 1/16 sample of the cube plus the full grey axis (full cube = GPU test)
    sample = cube[::4, ::4]
    assert np.array_equal(cx.bgr2hsv(sample), hsv[::4, ::4])
    assert np.array_equal(cx.bgr2lab(sample), lab[::4, ::4])


@pytest.mark.parametrize("bins,hi", [(18, 180), (8, 256)])
def test_calc_hist(bins, hi):
    rng = np.random.default_rng(0)
    ch = rng.integers(0, hi, (77, 55)).astype(np.uint8)
    ref = cv2.calcHist([ch], [0], None, [bins], [0, hi]).flatten()
    assert ref.dtype == np.float32
    assert np.array_equal(ref, cx.calc_hist_u8(ch, bins, hi))


PIL_SHAPES = [(1, 1), (5, 3), (40, 20), (75, 33), (125, 66), (250, 110), (128, 64), (200, 64), (128, 100),
              (300, 2), (201, 2), (200, 2), (401, 4), (64, 128), (17, 200), (500, 300), (101, 1), (100, 1)]


@pytest.mark.parametrize("h,w", PIL_SHAPES)
def test_pil_resize_bit_exact(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)):
        ref = np.asarray(Image.fromarray(img).resize((64, 128), Image.BILINEAR))
        assert np.array_equal(ref, cx.pil_resize_bilinear(img, 64, 128))


def test_mnv3_preprocess_matches_torchvision():
    from oracle.team_reference import make_preprocess
    rng = np.random.default_rng(3)
    pp = make_preprocess()
    for (h, w) in ((125, 66), (40, 20), (7, 9), (300, 2)):
        roi = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(pp(roi).numpy(), cx.mnv3_preprocess(roi))


CV_CASES = [(1080, 1920, 720, 1280), (2160, 3840, 720, 1280), (112, 256, 280, 640), (624, 256, 640, 263),
            (720, 1280, 360, 640), (1, 1, 5, 7), (2, 2, 9, 5), (33, 47, 64, 64), (100, 50, 50, 100), (7, 9, 7, 18),
            (480, 640, 240, 640), (300, 300, 100, 100)]


@pytest.mark.parametrize("sh,sw,dh,dw", CV_CASES)
def test_cv_resize_bit_exact(sh, sw, dh, dw):
    rng = np.random.default_rng(sh + dw)
    for img in (rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8), (rng.integers(0, 2, (sh, sw, 3)) * 255).astype(np.uint8)):
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, cx.cv_resize_linear(img, dw, dh))
