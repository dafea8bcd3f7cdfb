"""Launch the K1 letterbox kernel a few times on synthetic frames (for ncu / quick timing)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))
from hvb import _ffi
from hvb.runtime import get_context
ctx = get_context(0)
rng = np.random.default_rng(0)
def run(n, h, w, mode, imgsz, reps=20, tag=""):
    frames = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
    plan = ctx.letterbox_plan(n, h, w, mode, imgsz)
    out = plan.run(frames)
    torch.cuda.synchronize()
    # many back-to-back launches between ONE event pair: the GPU stays busy, so host launch gaps
    # (python/ctypes overhead) are not counted as kernel time
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        plan.run(frames, out)
    b.record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) / reps])
    gb = (plan.read_bytes + plan.write_bytes) / 1e9
    print("%s n=%d %dx%d mode=%d: %.1f us  %.0f GB/s (min %.1f us)" % (tag, n, w, h, mode, 1e3 * ms.mean(), gb / (ms.mean() / 1e3), 1e3 * ms.min()))
run(16, 1080, 1920, _ffi.LB_WHOLE, 1280, tag="K1a")
run(4, 2160, 3840, _ffi.LB_SLICE_EXACT, 640, tag="K1b")
run(32, 1080, 1920, _ffi.LB_WHOLE, 1280, tag="K1a")
run(8, 2160, 3840, _ffi.LB_SLICE_UNIFORM, 640, tag="K1b-uniform")
