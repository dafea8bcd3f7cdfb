"""Probe: cost of staging a 64-frame 1080p chunk (398 MB) into the pinned buffer — torch copy_ per frame from the main
thread / a side thread, numpy copyto, and a thread pool of GIL-releasing copies.  python tools/probe_staging.py"""
import json, os, sys, threading, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hockey-vision-analytics_b200")]
import torch

n, h, w = 64, 1080, 1920
frames = [np.random.default_rng(i).integers(0, 255, (h, w, 3), dtype=np.uint8) for i in range(n)]
buf = torch.empty((n, h, w, 3), dtype=torch.uint8)
if torch.cuda.is_available():
    buf = buf.pin_memory()
bufn = buf.numpy()
out = {"torch_threads": torch.get_num_threads(), "cpus": os.cpu_count()}


def t_torch():
    for k, f in enumerate(frames):
        buf[k].copy_(torch.from_numpy(f))


def t_numpy():
    for k, f in enumerate(frames):
        np.copyto(bufn[k], f)


def timed(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / reps


def in_thread(fn):
    res = {}
    th = threading.Thread(target=lambda: res.setdefault("ms", timed(fn)))
    th.start(); th.join()
    return res["ms"]


out["torch_copy_main_ms"] = timed(t_torch)
out["numpy_copyto_main_ms"] = timed(t_numpy)
out["torch_copy_side_thread_ms"] = in_thread(t_torch)
out["numpy_copyto_side_thread_ms"] = in_thread(t_numpy)
for T in (2, 4, 8):
    pool = ThreadPoolExecutor(T)

    def pooled():
        def part(lo):
            for k in range(lo, n, T):
                np.copyto(bufn[k], frames[k])
        list(pool.map(part, range(T)))
    out["numpy_pool%d_ms" % T] = timed(pooled)
    out["numpy_pool%d_side_thread_ms" % T] = in_thread(pooled)
    pool.shutdown()
import ctypes
from hvb import _ffi
ptrs = (ctypes.c_void_p * n)(*[f.ctypes.data for f in frames])
for T in (1, 2, 4, 8):
    fn = lambda: _ffi.check(_ffi.lib().hvb_stage_frames(ctypes.cast(ptrs, ctypes.c_void_p), n, h * w * 3, buf.data_ptr(), T))
    out["hvb_stage_frames_%d_threads_ms" % T] = timed(fn)
    out["hvb_stage_frames_%d_threads_side_thread_ms" % T] = in_thread(fn)
torch.set_num_threads(1)
out["torch_copy_1thread_main_ms"] = timed(t_torch)
print(json.dumps(out))
