"""Generate tests/golden/segmentation_reference.npz by running the REAL reference
SegmentationTeamClassifier (hockey/common/team_segmentation.py, imported from /root/reference in the build
container) on the seeded crops of make_golden.golden_crops().

    python tests/golden/make_golden_segmentation.py

GrabCut is out of scope (SURVEY.md §8f rank 4 asks for the colour features without it), so ``cv2.grabCut`` is
patched to raise: the reference's own ``except`` branch (team_segmentation.py:87-96) then returns its fallback
rectangle, and everything downstream (extract_jersey_colors, classify_single_jersey, fit, predict) is the
reference's unmodified code.  Only OUTPUTS are stored; inputs are regenerated from seeds.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import golden_crops  # noqa: E402


def load_real():
    from oracle import reference_loader as rl
    assert rl.available(), "run this in the build container where /root/reference is mounted"
    rl.load(0)                                   # installs the supervision stub and the sys.path entry
    import cv2
    import common.team_segmentation as seg       # noqa: E402  (reference module)

    def no_grabcut(*a, **k):
        raise cv2.error("GrabCut disabled: fallback rectangle requested")

    seg.cv2.grabCut = no_grabcut
    return seg


def run_real(seg, crops, tids, n_fit):
    out = {}
    clf = seg.SegmentationTeamClassifier()
    masks = [clf.segment_player(c) for c in crops]
    out["mask_rect"] = np.array([[np.argmax(m.any(1)) if m.any() else 0, m.any(1).sum(), np.argmax(m.any(0)) if m.any() else 0,
                                  m.any(0).sum(), m.sum()] for m in masks])
    feats = [clf.extract_jersey_colors(c, m) for c, m in zip(crops, masks)]
    out["features"] = np.array([[f["is_white"], f["dominant_hue"], f["saturation"], f["brightness"]] for f in feats], np.float64)
    cj = [clf.classify_single_jersey(c) for c in crops]
    out["single_team"] = np.array([t for t, _ in cj])
    out["single_conf"] = np.array([c for _, c in cj], np.float64)
    out["predict_unfitted"] = clf.predict(list(crops[:n_fit]), tids[:n_fit])
    clf2 = seg.SegmentationTeamClassifier()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf2.fit(list(crops[:n_fit]))
    out["centers"] = clf2.kmeans.cluster_centers_
    per = n_fit // 4
    out["predict"] = np.concatenate([clf2.predict(list(crops[f * per:(f + 1) * per]), tids[f * per:(f + 1) * per]) for f in range(4)])
    out["predict_no_ids"] = clf2.predict(list(crops[:n_fit]))
    return out


def main():
    seg = load_real()
    _, crops, labels, _, tids = golden_crops()
    out = run_real(seg, crops, tids, int((labels >= 0).sum()))
    np.savez_compressed(os.path.join(HERE, "segmentation_reference.npz"), **out)
    print("wrote segmentation_reference.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
