#!/usr/bin/env python
"""bench.py — hot-path frames/s on synthetic 1080p frames (+ the 4K sliced puck path), with the
roofline of the dominant libhvb kernel and the CPU reference path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl hvb|reference] [--chunk F]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one chunk of F (default 64) synthetic 1080p frames per GPU
(BASELINE.json configs[1] + configs[2]):
    K1a letterbox -> YOLOv8m forward (torch fp32, random init) -> K2a decode+NMS ->
    K3a colour features + K3b crop preprocessing on 12 player boxes / frame -> MobileNetV3 (torch fp32) ->
    K4a scale_transform -> similarity rule.
Random-init YOLO emits no detection above conf=0.4, so the team stage runs on the
12 planted player boxes of each synthetic frame ("team_boxes": "planted"); the detection stage
still does all of its work.  `value` has the frames resident in HBM; `e2e` goes through the public
host API (pinned host frames in, H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "hockey-vision-analytics_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W, PLAYERS = 1080, 1920, 12
METRIC = "hot_path_frames_per_sec_1080p"
UNIT = "frames/s"


def ncu_traffic(kernel_key: str, launches=None):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the roofline kernel from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json, written from tools/ncu_summary.py output).
    `launches`: only use a per-step capture if it was taken with the same number of launches per step as now."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        for key in (kernel_key, kernel_key + "_k6"):      # "_k6": the capture minus the launches K6 now replaces
            e = d.get(key, {})
            if e and (launches is None or e.get("launches") == launches):
                return e.get("traffic_bytes")
        return None
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_chunk(seed: int, n: int):
    from hvb.synth import rink_frame
    rng = np.random.default_rng(seed)
    frames, boxes, fidx = [], [], []
    for i in range(n):
        f, b, _, _ = rink_frame(rng, H, W, PLAYERS)
        frames.append(f); boxes.append(b); fidx.append(np.full(len(b), i, np.int32))
    return np.stack(frames), np.concatenate(boxes).astype(np.float32), np.concatenate(fidx)


# ------------------------------------------------------------------------------------------ reference arm
class CpuReferencePath:
    """The reference's own CPU implementation of the path, restated in oracle/ against the same real
    libraries (cv2, Pillow/torchvision, sklearn, torchvision.ops.nms): kind = "port"."""

    def __init__(self):
        import torch
        from hvb.models import build_trunk, build_yolov8
        from oracle import team_reference as tr
        self.torch = torch
        from hvb.models.yolov8 import fuse_conv_bn
        self.yolo = fuse_conv_bn(build_yolov8("m", 2, 0))        # ultralytics fuses conv+bn before inference
        self.trunk = build_trunk(0, calibrate=True)
        self.ref = tr.HybridReference(self.trunk)
        self.fitted = False
        self.stage_s = {"letterbox+preprocess": 0.0, "yolo_forward": 0.0, "decode+nms+scale": 0.0, "crops+team_predict": 0.0}
        self.stage_frames = 0

    def fit(self, frames, boxes, fidx):
        from oracle.supervision_restated import crop_image
        crops = [crop_image(frames[f], b) for f, b in zip(fidx, boxes)]
        self.ref.fit(crops, run_clustering=False)
        self.fitted = True

    def step(self, frames, boxes, fidx):
        from oracle import ultralytics_restated as ur
        from oracle.supervision_restated import crop_image
        torch = self.torch
        out = []
        st = self.stage_s
        for i, frame in enumerate(frames):
            t0 = time.perf_counter()
            lb = ur.letterbox(frame, 1280, auto=True)
            x = torch.from_numpy(ur.preprocess([lb]))
            t1 = time.perf_counter()
            with torch.no_grad():
                heads = self.yolo(x)
            t2 = time.perf_counter()
            det = ur.predict_from_head(heads, 2, tuple(x.shape[2:]), [frame.shape[:2]], 0.4)[0]
            t3 = time.perf_counter()
            crops = [crop_image(frame, b) for b in boxes[fidx == i]]
            out.append((det, self.ref.predict(crops)))
            t4 = time.perf_counter()
            st["letterbox+preprocess"] += t1 - t0; st["yolo_forward"] += t2 - t1
            st["decode+nms+scale"] += t3 - t2; st["crops+team_predict"] += t4 - t3
            self.stage_frames += 1
        return out

    def stage_ms_per_frame(self):
        n = max(self.stage_frames, 1)
        return {k: round(1e3 * v / n, 3) for k, v in self.stage_s.items()}


def run_reference(args, rank, world):
    import torch
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    path = CpuReferencePath()
    frames, boxes, fidx = synth_chunk(0, args.ref_frames)
    path.fit(frames, boxes, fidx)
    for _ in range(args.warmup):
        path.step(frames[:1], boxes[fidx == 0], fidx[fidx == 0])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        path.step(frames, boxes, fidx)
    dt = time.perf_counter() - t0
    fps = args.steps * len(frames) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": "C2+C3 1080p player detection + team classification, CPU reference path",
                   "frames_per_step": int(len(frames)), "players_per_frame": PLAYERS, "team_boxes": "planted"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "stage_ms_per_frame": path.stage_ms_per_frame(),
                         "sample": "%d synthetic 1080p frames per step: cv2 letterbox + YOLOv8m CPU forward + restated "
                                   "decode/torchvision NMS + reference colour/MobileNetV3 features + predict" % len(frames)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ hvb arm
def run_hvb(args, rank, world):
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    from hvb import _ffi
    from hvb.pipeline import HotPath, SlicedPuckPath
    from hvb.runtime import get_context

    ctx = get_context(local)
    F = args.chunk
    path = HotPath(dev, "m", 2, 1280, 0.4, seed=0)
    frames, boxes, fidx = synth_chunk(1000 + rank, F)             # each rank owns its own clip chunk (weak scaling)
    pinned = torch.from_numpy(frames).pin_memory()
    frames_dev = pinned.to(dev)
    boxes_dev = torch.from_numpy(boxes).to(dev)
    fidx_dev = torch.from_numpy(fidx).to(dev)

    # ---- one-off team fit (per job): local crop features -> all-gather (NCCL) -> global standardise + affinity
    t_fit0 = time.perf_counter()
    feats, raw, _ = path.classifier.features_from_frame(frames_dev, boxes_dev, fidx_dev)
    if world > 1:
        from hvb.dist import all_gather_features
        feats = all_gather_features(feats)
    # scaler + RBF affinity on the device on every rank; the spectral embedding / k-means that follow in the
    # reference's fit run in scikit-learn on the host, are not on the per-frame path and are skipped here
    path.classifier.fit_features(feats, None, None, cluster=False)
    torch.cuda.synchronize()
    fit_ms = 1e3 * (time.perf_counter() - t_fit0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            torch.cuda.profiler.start()          # ncu --profile-from-start off: only the timed steps are captured
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if profile:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    # ---- (i) device-resident hot path
    k1_events = []

    def step_device():
        path.process_chunk_device(frames_dev, boxes_dev, fidx_dev)

    ctx.launch_count(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, clocks = timed(step_device, args.steps, args.warmup, sampler, profile=args.profile_region and not args.profile_4k)
    launches = ctx.launch_count() // (args.steps + args.warmup) * args.steps
    fps_dev = world * F * args.steps / (ms_dev / 1e3)

    # ---- (ii) end to end through the public host API (pinned frames, H2D + D2H inside)
    # Every step copies its own chunk from pinned host memory and reads its results back; the pipelined
    # API (HotPath.process_stream) overlaps the copy of step i+1 with the kernels of step i.
    def run_e2e(k):
        n = 0
        for res in path.process_stream((pinned, boxes, fidx) for _ in range(k)):
            n += len(res["team"])
        return n

    run_e2e(max(2, args.warmup // 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    fps_e2e = world * F * args.steps / (ms_e2e / 1e3)
    h2d = frames.nbytes + boxes.nbytes + fidx.nbytes
    md = path.detector.max_det
    d2h = F * md * (16 + 4 + 4) + F * 4 + len(boxes) * 80

    # ---- roofline of the dominant libhvb kernel (K1 letterbox), timed live with CUDA events on its stream
    plan = path.detector.plan(F, H, W, _ffi.LB_WHOLE)
    out = plan.run(frames_dev)
    torch.cuda.synchronize()
    # `reps` back-to-back launches between one pair of CUDA events on the launching stream: the GPU
    # stays busy, so host launch gaps are not billed to the kernel; inputs+outputs exceed L2.
    reps = max(args.steps, 20)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(reps):
        plan.run(frames_dev, out)
    eb.record()
    torch.cuda.synchronize()
    k1_ms = ea.elapsed_time(eb) / reps
    k1_bytes = plan.read_bytes + plan.write_bytes
    peak, peak_src = measured_peaks()
    achieved = k1_bytes / (k1_ms / 1e3) / 1e9

    # ---- roofline of the DOMINANT libhvb kernel of the step: the K5 conv epilogue (bias_act_kernel, ~28 % of the step,
    # 82 launches, 76 since K6 took six of them together with their convolutions).  Every launch of three more steps is bracketed by CUDA events on its launching stream.
    runner = path.detector.runner
    k5 = None
    if runner is not None:
        runner.epi_log = []
        for _ in range(3):
            path.detect_device(frames_dev)
        torch.cuda.synchronize()
        log, runner.epi_log = runner.epi_log, None
        k5_bytes = sum(b for b, _, _ in log)
        k5_ms = sum(a.elapsed_time(b) for _, a, b in log)
        k5 = {"launches_per_step": len(log) // 3, "bytes_per_step": k5_bytes // 3, "ms_per_step": k5_ms / 3,
              "achieved": k5_bytes / (k5_ms / 1e3) / 1e9}

    # ---- per-stage device times of one chunk (CUDA events on the main stream, stages run back to back without the
    # side-stream overlap) — the GPU column next to cpu_baseline.stage_ms_per_frame
    det = path.detector
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = [0.0, 0.0, 0.0, 0.0]
    for it in range(4):
        plan = det.plan(F, H, W, _ffi.LB_WHOLE)
        ev[0].record()
        x = plan.class_views(plan.run(frames_dev))[0]
        ev[1].record()
        heads = det.forward_heads(x)
        ev[2].record()
        meta_h, meta_d = det._meta_dev(plan, 0)
        det._decode(heads, meta_h, meta_d, F)
        ev[3].record()
        path.team_device(frames_dev, boxes_dev, fidx_dev)
        ev[4].record()
        torch.cuda.synchronize()
        if it:                                     # first pass is a warm-up
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) / 3 / F
    gpu_stage_ms = {"letterbox+preprocess (K1a)": round(acc[0], 5), "yolo_forward (cuDNN convs + K5)": round(acc[1], 5),
                    "decode+nms+scale (K2a)": round(acc[2], 5), "crops+team_predict (K3a/K3b, MobileNetV3, K4a)": round(acc[3], 5)}

    # ---- secondary workload: 4K sliced puck path (C4), reported in `extra`
    extra = {"fit_ms": fit_ms, "gpu_stage_ms_per_frame": gpu_stage_ms}
    # frame-at-a-time use of the drop-in (what process_frame does per frame: host frame in, Detections out)
    if rank == 0:
        one = frames[0]
        for tag, flag in (("eager", False), ("cuda_graph", True)):
            path.detector.cuda_graph = flag
            for _ in range(3):
                path.detector.detect_players(one)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                path.detector.detect_players(one)
            extra["frame_at_a_time_detect_fps_" + tag] = 20 / (time.perf_counter() - t0)
        path.detector.cuda_graph = False
        # the whole-loop drop-in on a clip (hvb.VideoProcessor.process_video_chunked: detection per chunk, ByteTrack per
        # frame on the host, team features per chunk).  Random-init YOLO detects nothing above conf 0.4, so this times the
        # driver + detection path; the tracker / team stages see empty inputs.
        from hvb.video import Config, VideoProcessor
        from hvb.models import build_yolov8
        vp = VideoProcessor(build_yolov8("m", 2, 0), dev, Config(), team_classifier=path.classifier_router())
        CH = 32                                                       # process_video_chunked's default chunk
        clip = [frames[i % F] for i in range(6 * CH)]
        list(vp.process_video_chunked(clip[:CH], chunk=CH, initialize=False))
        t0 = time.perf_counter()
        n_out = len(list(vp.process_video_chunked(clip, chunk=CH, initialize=False)))
        extra["clip_chunked_drop_in_fps"] = n_out / (time.perf_counter() - t0)
        del vp
    if args.with_4k:
        from hvb.synth import rink_frame
        rng = np.random.default_rng(7 + rank)
        F4 = args.chunk_4k
        f4 = np.stack([rink_frame(rng, 2160, 3840, PLAYERS, 2.0)[0] for _ in range(F4)])
        f4_dev = torch.from_numpy(f4).to(dev)
        puck = SlicedPuckPath(dev, "n", 1, 0.4)
        ms4_eager, _ = timed(lambda: puck.process_chunk_device(f4_dev), max(2, args.steps // 2), 2,
                             profile=args.profile_region and args.profile_4k)
        # the sliced path is launch-bound (≈900 launches per chunk over 5 tile shape classes): replay it as one CUDA graph
        ms4, _ = timed(lambda: puck.process_chunk_device(f4_dev, graph=True), max(2, args.steps // 2), 2)
        # end to end through the public sliced API: pinned 4K frames in, per-frame Detections out
        pinned4 = torch.from_numpy(f4).pin_memory()
        k4 = max(2, args.steps // 2)

        def run_e2e4(k):
            return sum(len(r) for r in puck.process_stream(pinned4 for _ in range(k)))

        run_e2e4(2)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        run_e2e4(k4)
        eb.record()
        barrier()
        ms4_e2e = ea.elapsed_time(eb)
        if world > 1:
            t = torch.tensor([ms4_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms4_e2e = float(t.item())
        extra["c4_4k_sliced_e2e_frames_per_sec"] = world * F4 * k4 / (ms4_e2e / 1e3)
        extra["c4_e2e_h2d_bytes_per_step"] = int(f4.nbytes)
        plan4 = puck.detector.plan(F4, 2160, 3840, _ffi.LB_SLICE_EXACT, 640, (640, 640), (128, 128))
        o4 = plan4.run(f4_dev)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(20):
            plan4.run(f4_dev, o4)
        eb.record()
        torch.cuda.synchronize()
        k1b_ms = ea.elapsed_time(eb) / 20
        extra.update({"c4_4k_sliced_frames_per_sec": world * F4 * max(2, args.steps // 2) / (ms4 / 1e3),
                      "c4_4k_sliced_frames_per_sec_eager_launches": world * F4 * max(2, args.steps // 2) / (ms4_eager / 1e3),
                      "c4_mode": "whole chunk replayed as one CUDA graph",
                      "c4_frames_per_step": F4, "c4_tiles_per_frame": int(plan4.tiles_per_frame),
                      "k1b_slice_letterbox_gbs": (plan4.read_bytes + plan4.write_bytes) / (k1b_ms / 1e3) / 1e9,
                      "k1b_frac_of_hbm_peak": (plan4.read_bytes + plan4.write_bytes) / (k1b_ms / 1e3) / 1e9 / peak})

    # ---- CPU baseline beside it (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref = CpuReferencePath()
        nf = args.ref_frames
        ref.fit(frames[:nf], boxes[fidx < nf], fidx[fidx < nf])
        ref.step(frames[:1], boxes[fidx == 0], fidx[fidx == 0])
        ref.stage_s = {k: 0.0 for k in ref.stage_s}; ref.stage_frames = 0
        t0 = time.perf_counter()
        ref.step(frames[:nf], boxes[fidx < nf], fidx[fidx < nf])
        dt = time.perf_counter() - t0
        stage_ms = ref.stage_ms_per_frame()
        # per-core normalisation: the same path on ONE host thread, one frame
        import cv2
        cv_threads = cv2.getNumThreads()
        torch.set_num_threads(1); cv2.setNumThreads(1)
        t0 = time.perf_counter()
        ref.step(frames[:1], boxes[fidx == 0], fidx[fidx == 0])
        dt1 = time.perf_counter() - t0
        torch.set_num_threads(cores); cv2.setNumThreads(cv_threads)
        cpu = {"value": nf / dt, "unit": UNIT, "cores": cores, "kind": "port", "torch_threads": cores, "cv2_threads": cv_threads,
               "value_1_thread": 1.0 / dt1, "stage_ms_per_frame": stage_ms,
               "sample": "%d of the %d synthetic 1080p frames of one step, same stages on the host: cv2 letterbox, YOLOv8m "
                         "CPU forward (torch, %d threads), restated decode + real torchvision NMS, reference colour + "
                         "MobileNetV3 features, predict" % (nf, F, cores)}

    k1_roof = {"kernel": "letterbox_kernel<false> (K1a, 1080p->736x1280, %d frames/launch)" % F, "bound": "hbm",
               "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
               "traffic": ncu_traffic("K1a_1080p_x%d" % F), "algorithmic_bytes_per_launch": int(k1_bytes), "avg_launch_ms": k1_ms}
    if k5 is not None:
        n5 = k5["launches_per_step"]
        t5 = ncu_traffic("K5_bias_act_step_x%d" % F, launches=n5)   # summed over the launches of one step
        roofline = {"kernel": "bias_act_kernel (K5 conv epilogue: bias + SiLU + residual -> dense / concat-slice destinations), "
                              "%d launches per step of %d frames" % (n5, F),
                    "bound": "hbm", "achieved": k5["achieved"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": k5["achieved"] / peak, "traffic": (t5 / n5) if t5 else None,
                    "algorithmic_bytes_per_launch": k5["bytes_per_step"] / n5, "avg_launch_ms": k5["ms_per_step"] / n5,
                    "share_of_step": k5["ms_per_step"] / (ms_dev / args.steps)}
        extra["roofline_k1a"] = k1_roof
    else:
        roofline = k1_roof
    if rank == 0:
        line = {
            "metric": METRIC, "value": fps_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": "C2+C3: 1080p player detection (K1a letterbox, YOLOv8m fp32 random-init, K2a decode+NMS) + "
                                   "team classification (K3a/K3b on 12 planted player boxes per frame, MobileNetV3-small fp32, "
                                   "K4a scale_transform, rule)",
                       "frames_per_step_per_gpu": F, "players_per_frame": PLAYERS, "team_boxes": "planted",
                       "backbones": "convolutions in torch/cuDNN (fp32 storage, conv+bn folded, channels_last, cudnn.benchmark, TF32 for YOLO / "
                                    "TF32 off for MobileNetV3); everything between the YOLO convolutions (bias, SiLU, residual, concat, "
                                    "upsample, layer 0) in libhvb K5 kernels; the pointwise convolutions with few channels as single K6 launches (tcgen05 TF32 GEMM + epilogue)",
                       "l2": "inputs larger than L2 (%.0f MB frames + %.0f MB letterboxed per step)" % (frames.nbytes / 1e6, k1_bytes / 1e6),
                       "parallelism": "frame chunks sharded per GPU, no data-path collective; one NCCL feature all-gather at fit"},
            "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hvb", choices=["hvb", "reference"])
    ap.add_argument("--chunk", type=int, default=64, help="1080p frames per step per GPU (64 measured best: 1832 vs 1760 frames/s at 32, 1734 at 48, 1753 at 96)")
    ap.add_argument("--chunk-4k", type=int, default=16, help="4K frames per step of the sliced puck path (640 tiles per step)")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per CPU-reference step (bounded sample)")
    ap.add_argument("--no-4k", dest="with_4k", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-4k", action="store_true", help="with --profile-region: bracket the eager 4K sliced steps instead of the 1080p steps")
    ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart/Stop around the timed device steps (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "hvb" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        run_hvb(args, rank, world)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
