"""Generate tests/golden/team_*.npz by running the REAL reference (imported from /root/reference in
the build container, see oracle/reference_loader.py) on seeded synthetic crops.

    python tests/golden/make_golden.py

The inputs are regenerated from seeds by `golden_crops()` (shared with the tests), so only the
reference's OUTPUTS are stored.  /root/reference does not exist on the GPU box; the committed
.npz files are what travels.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "hockey-vision-analytics_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def golden_crops(seed: int = 7, n_frames: int = 4, n_players: int = 10):
    """Crops as the reference's main loop produces them: NON-contiguous views into 1080p frames
    (sv.crop_image at hockey/main.py:326), plus a few degenerate sizes."""
    from hvb.synth import rink_clip
    from oracle.supervision_restated import crop_image
    frames, boxes, teams, _ = rink_clip(seed, n_frames, 1080, 1920, n_players)
    crops, labels, positions, tids = [], [], [], []
    for f in range(n_frames):
        for i, b in enumerate(boxes[f]):
            jit = np.array([0.4, -0.3, 0.6, 0.2], np.float32) * (i % 3)
            crops.append(crop_image(frames[f], b + jit))
            labels.append(int(teams[f][i]))
            positions.append((float((b[0] + b[2]) / 2), float((b[1] + b[3]) / 2)))
            tids.append(i + 1)
    rng = np.random.default_rng(seed + 1)
    for (h, w) in ((1, 1), (39, 25), (40, 19), (40, 20), (64, 33), (150, 70), (300, 17)):
        crops.append(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
        labels.append(-1)
        positions.append((0.0, 0.0))
        tids.append(100 + h)
    return frames, crops, np.asarray(labels), positions, np.asarray(tids)


def main():
    from oracle import reference_loader as rl
    assert rl.available(), "run this in the build container where /root/reference is mounted"
    import torch
    torch.set_num_threads(1)
    hyb = rl.make_hybrid(seed=0)
    _, team_mod = rl.load(0)
    frames, crops, labels, positions, tids = golden_crops()

    out = {}
    out["color"] = hyb.extract_color_features(crops)
    out["deep"] = hyb.extract_deep_features(crops).astype(np.float32)
    out["preprocessed"] = np.stack([hyb.preprocess(hyb.extract_jersey_region(c)).numpy() for c in crops])
    out["jersey_shapes"] = np.array([hyb.extract_jersey_region(c).shape[:2] for c in crops])

    # fit on the player crops (positions omitted, as TeamClassifier.fit does -> D = 625)
    n_fit = int((labels >= 0).sum())
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hyb.fit(crops[:n_fit])
    out["scaler_mean"], out["scaler_scale"] = hyb.scaler.mean_, hyb.scaler.scale_
    out["affinity"] = hyb.clusterer.affinity_matrix_
    out["cluster_labels"] = hyb.cluster_labels
    feats = hyb.extract_all_features(crops[:n_fit])
    out["features_scaled"] = hyb.scaler.transform(feats)
    # predict frame by frame with tracker ids (exercises the temporal vote)
    per = n_fit // 4
    preds = []
    for f in range(4):
        sl = slice(f * per, (f + 1) * per)
        preds.append(hyb.predict(crops[sl], tids[sl]))
    out["predict"] = np.concatenate(preds)
    out["predict_no_ids"] = hyb.predict(crops[:n_fit])

    # 627-d variant: fit with positions called directly on the hybrid class
    hyb2 = rl.make_hybrid(seed=0)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hyb2.fit(crops[:n_fit], positions[:n_fit])
    out["affinity_pos"] = hyb2.clusterer.affinity_matrix_

    # simple HSV rule of team.py
    with contextlib.redirect_stdout(io.StringIO()):
        simple = team_mod.TeamClassifier(use_hybrid=False, use_robust=False, use_interactive=False, use_segmentation=False)
    cj = [simple.classify_jersey(c) for c in crops]
    out["simple_team"] = np.array([t for t, _ in cj])
    out["simple_conf"] = np.array([c for _, c in cj], np.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        out["simple_predict"] = simple.predict(crops[:n_fit], tids[:n_fit])

    np.savez_compressed(os.path.join(HERE, "team_reference.npz"), **out)
    print("wrote team_reference.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
