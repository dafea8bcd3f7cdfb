"""Probe 3: K1a inside a process that has built the detector and run the chunk pipeline (bench.py's situation)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hockey-vision-analytics_b200")]
import torch
from hvb import _ffi
from hvb.pipeline import HotPath
from hvb.synth import rink_clip
from hvb.video import Config, VideoProcessor

dev = "cuda:0"
F, H, W = 64, 1080, 1920
res = {}


def timed(plan, frames, out, reps=20):
    plan.run(frames, out); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        plan.run(frames, out)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return {"us": round(1e3 * a.elapsed_time(b) / reps, 1), "host_issue_us": round(1e6 * (t1 - t0) / reps, 1)}


frames = torch.from_numpy(rink_clip(1000, F, H, W, 12, 1.0)[0]).to(dev)
from hvb.runtime import get_context
ctx = get_context(0)
plan0 = ctx.letterbox_plan(F, H, W, _ffi.LB_WHOLE, 1280)
out0 = plan0.run(frames)
res["before anything else"] = timed(plan0, frames, out0)
path = HotPath(dev, "m", 2, 1280, 0.4, seed=0)
det = path.detector
res["after building HotPath"] = timed(plan0, frames, out0)
plan1 = det.plan(F, H, W, _ffi.LB_WHOLE)
res["det.plan (same geometry)"] = timed(plan1, frames, out0)
for _ in range(3):
    path.detect_device(frames)
torch.cuda.synchronize()
res["after 3 YOLO forwards (cudnn.benchmark on)"] = timed(plan0, frames, out0)
out1 = plan1.run(frames)
res["fresh output tensor after the forwards"] = timed(plan1, frames, out1)
res["cudnn.benchmark"] = torch.backends.cudnn.benchmark
torch.cuda.empty_cache()
out2 = plan1.run(frames)
res["after empty_cache, new output"] = timed(plan1, frames, out2)
print(json.dumps(res))
