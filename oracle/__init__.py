"""CPU oracle for the hvb hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of what the reference
(JetJadeja/hockey-vision-analytics) computes on its per-frame hot path, either
directly (``hockey/common/team_hybrid.py``, ``team.py``, ``main.py``) or through
the third-party libraries it calls (OpenCV, Pillow, torchvision, scikit-learn,
ultralytics 8.3.148, supervision >= 0.21).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker.
The product (``hockey-vision-analytics_b200/``) never imports it and fails
loudly when ``libhvb.so`` is missing.

Pinning status (the reference ships no tests / golden vectors):
  * cv_exact.*            pinned against the installed cv2 4.13 / Pillow 12.2
                          (full 2^24 colour cube sha256, random images).
  * team_reference.*      pinned against the reference's own team_hybrid.py /
                          team.py imported in the build container
                          (tests/golden/team_*.npz, made by tests/golden/make_golden.py).
  * ultralytics_restated  parity unpinned (ultralytics absent; restated from the
                          8.3.148 sources named in SURVEY.md App. B1; NMS uses the
                          real torchvision.ops.nms).
  * supervision_restated  parity unpinned (supervision absent; App. B2).
"""
