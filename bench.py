#!/usr/bin/env python
"""bench.py — hot-path frames/s on a synthetic 1080p clip (+ the 4K sliced puck path and the 720p config), with the
roofline of the dominant libhvb kernel and the CPU reference path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl hvb|reference] [--chunk F]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the reference's per-frame loop (hockey/main.py:259-281) over one chunk of F (default 64)
consecutive frames of a synthetic 1080p clip per GPU (BASELINE.json configs[1] + configs[2]):
    K1a letterbox -> YOLOv8m forward (random init) -> planted candidates scattered into the raw head tensors ->
    K2a decode + NMS + scale_boxes -> K7 ByteTrack (every frame of the chunk, in order) -> crops of the tracked players ->
    K3a colour features + K3b crop preprocessing -> MobileNetV3 -> K4a scale_transform -> similarity rule + temporal vote
    -> per-frame results (detections, tracker ids, team ids, labels).
Random-init YOLO emits no detection above conf 0.4, so AFTER the real forward both arms overwrite the head
tensors at 36 anchors per frame with the same planted values (hvb.synth.PlantedOverlay: the 12 players of the synthetic
frame, 3 jittered candidates each, distinct confidences in (0.45, 0.97); 11 skaters + 1 goalie): NMS suppresses 24 of
36 candidates per frame, ByteTrack tracks 12 objects, and the team stage runs on the 11 tracked skaters the detector
found ("team_boxes": "detected").  `value` has the clip resident in HBM; `e2e` is the drop-in's public call
(hvb.VideoProcessor.process_video_chunked: host frames in, H2D inside, per-frame results out).
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

H, W, PLAYERS, DUP, IMGSZ = 1080, 1920, 12, 3, 1280
METRIC = "hot_path_frames_per_sec_1080p"
UNIT = "frames/s"
WORKLOAD = ("C2+C3: 1080p player detection (letterbox, YOLOv8m fp32 random-init, %d planted candidates per frame scattered "
            "into the head tensors, decode + NMS) + ByteTrack + team classification of the tracked players (colour + "
            "MobileNetV3-small features, scale_transform, rule + temporal vote)" % (PLAYERS * DUP))


def ncu_traffic(kernel_key: str, launches=None):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the roofline kernel from the committed ncu capture
    (profiles/ncu_traffic.json).  `launches`: only use a per-step capture taken with the same number of launches per step."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        table = json.load(open(p))
        for key in ([kernel_key] if isinstance(kernel_key, str) else kernel_key):      # newest capture first
            e = table.get(key, {})
            if e and (launches is None or e.get("launches") == launches):
                return e.get("traffic_bytes")
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
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
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: nvidia-smi is started before the warm-up (it needs up to a second to emit its first
        line, longer than a short timed region), only samples taken after this mark are reported."""
        self.t_mark = time.time()

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
        t_mark = getattr(self, "t_mark", 0.0)
        inside = [l for t, l in self.lines if t >= t_mark]
        for l in inside or [l for _, l in self.lines[-2:]]:
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


def synth_clip(seed: int, n: int, h: int = H, w: int = W, scale: float = 1.0):
    """n consecutive frames of a synthetic clip (players drift <= 8 px / frame), their boxes and classes."""
    from hvb.synth import rink_clip
    frames, boxes, _, pucks = rink_clip(seed, n, h, w, PLAYERS, scale)
    cls = [np.array([0] * (PLAYERS - 1) + [1])] * n                   # 11 skaters + 1 goalie
    return frames, boxes, cls, pucks


def make_overlay(seed: int, boxes, cls):
    from hvb.synth import PlantedOverlay
    return PlantedOverlay.whole_frame(seed, (H, W), IMGSZ, boxes, cls, nc=2, dup=DUP)


def pingpong(n_frames: int, n_chunks: int, start: int = 0):
    """Frame ids of chunk start, start+1, ...: the clip forwards, then backwards, ... (continuous motion for the tracker)."""
    fwd = list(range(n_frames))
    return [fwd if (start + k) % 2 == 0 else fwd[::-1] for k in range(n_chunks)]


# ------------------------------------------------------------------------------------------ reference arm
class CpuReferencePath:
    """The reference's own CPU implementation of the path, restated in oracle/ against the same real libraries (cv2,
    Pillow/torchvision, sklearn, torchvision.ops.nms, scipy): kind = "port".  One frame at a time, like main.py:321."""
    STAGES = ("letterbox+preprocess", "yolo_forward", "decode+nms+scale", "bytetrack", "crops+team_predict")

    def __init__(self, overlay):
        import torch
        from hvb.models import build_trunk, build_yolov8
        from hvb.models.yolov8 import fuse_conv_bn
        from oracle import team_reference as tr
        from oracle.bytetrack_restated import ByteTrack
        self.torch = torch
        self.yolo = fuse_conv_bn(build_yolov8("m", 2, 0))        # ultralytics fuses conv+bn before inference
        self.trunk = build_trunk(0, calibrate=True)
        self.ref = tr.HybridReference(self.trunk)
        self.tracker = ByteTrack(0.25, 30, 0.8, 30, 2)            # main.py:162-168
        self.overlay = overlay
        self.stage_s = {k: 0.0 for k in self.STAGES}
        self.stage_frames = 0
        self.n_tracked = 0

    def fit(self, frames, boxes):
        from oracle.supervision_restated import crop_image
        crops = [crop_image(f, b) for f, bb in zip(frames, boxes) for b in bb[:PLAYERS - 1]]
        self.ref.fit(crops, run_clustering=False)

    def step(self, frames, frame_ids):
        from oracle import ultralytics_restated as ur
        from oracle.supervision_restated import crop_image
        torch, st, out = self.torch, self.stage_s, []
        for frame, fid in zip(frames, frame_ids):
            t0 = time.perf_counter()
            x = torch.from_numpy(ur.preprocess([ur.letterbox(frame, IMGSZ, auto=True)]))
            t1 = time.perf_counter()
            with torch.no_grad():
                heads = [t.clone() for t in self.yolo(x)]
            self.overlay.apply_host(heads, fid)                   # the same planted candidates the GPU arm scatters in
            t2 = time.perf_counter()
            xyxy, conf, cls = ur.predict_from_head(heads, 2, tuple(x.shape[2:]), [frame.shape[:2]], 0.4)[0]
            m = ((cls == 0) | (cls == 1)) & (conf > 0.4)          # main.py:189-193
            xyxy, conf, cls = xyxy[m], conf[m], cls[m]
            t3 = time.perf_counter()
            keep, ids = self.tracker.update_with_detections(xyxy, conf)
            t4 = time.perf_counter()
            pl = cls[keep] == 0
            crops = [crop_image(frame, b) for b in xyxy[keep][pl]]
            team = self.ref.predict(crops, tracker_ids=ids[pl])
            t5 = time.perf_counter()
            out.append((xyxy[keep], ids, team))
            for k, dt in zip(self.STAGES, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                st[k] += dt
            self.stage_frames += 1
            self.n_tracked += len(keep)
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
    nf = args.ref_frames
    frames, boxes, cls, _ = synth_clip(1000, nf + 2)
    path = CpuReferencePath(make_overlay(77, boxes, cls))
    path.fit(frames, boxes)
    # untimed: the first two frames of the clip confirm the tracks (minimum_consecutive_frames = 2), so that every
    # timed frame has tracked players to classify; further warm-up steps repeat frame 1
    path.step(frames[:2], [0, 1])
    for _ in range(args.warmup):
        path.step(frames[1:2], [1])
    path.stage_s = {k: 0.0 for k in path.STAGES}; path.stage_frames = 0; path.n_tracked = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        ids = list(range(2, nf + 2)) if k % 2 == 0 else list(range(2, nf + 2))[::-1]
        path.step([frames[i] for i in ids], ids)
    dt = time.perf_counter() - t0
    fps = args.steps * nf / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + " — CPU reference path, one frame at a time", "frames_per_step": nf,
                   "players_per_frame": PLAYERS, "planted_candidates_per_frame": PLAYERS * DUP, "team_boxes": "detected"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "stage_ms_per_frame": path.stage_ms_per_frame(),
                         "tracked_per_frame": path.n_tracked / max(path.stage_frames, 1),
                         "sample": "%d consecutive synthetic 1080p frames per step: cv2 letterbox + YOLOv8m CPU forward (torch, %d threads) + "
                                   "planted candidates + restated decode / torchvision NMS + restated ByteTrack (scipy) + reference "
                                   "colour / MobileNetV3 features + predict" % (nf, cores)},
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
    # nvidia-smi needs seconds to emit its first line on an 8-GPU box: started here, long before the timed region (only
    # the samples taken after sampler.mark() are reported)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # every rank gets its share of the host cores (all ranks claiming all cores made the host-side API
    # slower at N > 1); pin the process to a contiguous block when the affinity mask allows it
    cores = os.cpu_count() or 1
    per_rank = max(1, cores // max(world, 1))
    torch.set_num_threads(per_rank)
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            avail = sorted(os.sched_getaffinity(0))
            if len(avail) >= world:
                k = len(avail) // world
                os.sched_setaffinity(0, set(avail[local * k:(local + 1) * k]))
        except OSError:
            pass
    try:
        import cv2
        cv2.setNumThreads(per_rank)
    except Exception:
        pass
    from hvb import _ffi
    from hvb.pipeline import HotPath, SlicedPuckPath
    from hvb.runtime import get_context
    from hvb.video import Config, VideoProcessor

    ctx = get_context(local)
    F = args.chunk
    path = HotPath(dev, "m", 2, IMGSZ, 0.4, seed=0)
    det = path.detector
    frames, boxes, cls, _ = synth_clip(1000 + rank, F)                 # each rank owns its own clip (weak scaling)
    overlay = make_overlay(77 + rank, boxes, cls)
    fwd_dev = torch.from_numpy(frames).to(dev)
    rev_dev = fwd_dev.flip(0).contiguous()
    hook_chunks = overlay.to_device(dev, pingpong(F, 2))               # cyclic: forwards, backwards
    hook_frames = overlay.to_device(dev, [[0]])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- one-off team fit (per job): local crop features -> all-gather (NCCL) -> global standardise + affinity.
    # The spectral embedding / k-means that follow in the reference's fit run in scikit-learn on the host, are not on
    # the per-frame path and are skipped here.  Run twice: the first pass carries cuDNN autotuning / allocations.
    skaters = np.concatenate([b[:PLAYERS - 1] for b in boxes]).astype(np.float32)
    sk_fidx = np.repeat(np.arange(F, dtype=np.int32), PLAYERS - 1)
    sk_dev, skf_dev = torch.from_numpy(skaters).to(dev), torch.from_numpy(sk_fidx).to(dev)
    fit = {}
    for tag in ("cold", "warm"):
        barrier()
        t0 = time.perf_counter()
        feats, _, _ = path.classifier.features_from_frame(fwd_dev, sk_dev, skf_dev)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1:
            from hvb.dist import all_gather_features
            feats = all_gather_features(feats)
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        path.classifier.fit_features(feats, None, None, cluster=False)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        fit[tag] = {"features_ms": 1e3 * (t1 - t0), "all_gather_ms": 1e3 * (t2 - t1), "standardise+affinity_ms": 1e3 * (t3 - t2),
                    "total_ms": 1e3 * (t3 - t0), "rows": int(feats.shape[0])}
    # the same exchange through the C ABI's own entry points (hvb_allgather_counts / hvb_allgather_features), and the opt-in
    # device clustering (K8: subspace iteration + Lloyd runs) on the gathered rows at the non-underflowing bandwidth
    if world > 1:
        from hvb.dist import all_gather_features
        local = path.classifier.features_from_frame(fwd_dev, sk_dev, skf_dev)[0]
        for _ in range(2):
            g2 = all_gather_features(local, backend="hvb")
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        g2 = all_gather_features(local, backend="hvb")
        torch.cuda.synchronize()
        fit["warm"]["all_gather_c_abi_ms"] = 1e3 * (time.perf_counter() - t0)
        fit["warm"]["all_gather_c_abi_equals_torch"] = bool(torch.equal(g2, feats))
    from hvb.spectral import DeviceSpectralClustering
    xs = path.classifier.features_normalized_
    _, aff = ctx.gram_affinity(xs, 1.0 / xs.shape[1], 0, want_d2=False, want_a=True)
    for _ in range(2):
        sc = DeviceSpectralClustering(2, 10, 42)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lab = sc.fit_predict(aff)
        fit["warm"]["device_spectral_clustering_ms (gamma = 1/625)"] = 1e3 * (time.perf_counter() - t0)
    fit["warm"]["device_spectral_outer_rounds"] = int(sc.info_.get("outer_iterations", 0))
    fit["warm"]["device_spectral_cluster_sizes"] = [int((lab == 0).sum()), int((lab == 1).sum())]
    del aff

    vp = VideoProcessor(device=dev, config=Config(), detector=det, team_classifier=path.classifier_router(), tracker=args.tracker)
    det.head_hook = hook_chunks
    state = {"g": 0, "tracked": 0, "players": 0, "frames": 0}

    def run_chunks(n):
        """n more chunks of the ping-pong clip through the device-resident chunk pipeline."""
        g0 = state["g"]
        state["g"] += n
        for r in vp.process_chunks((fwd_dev if (g0 + k) % 2 == 0 else rev_dev) for k in range(n)):
            state["tracked"] += len(r.detections); state["players"] += len(r.player_team_ids); state["frames"] += 1

    # ---- (i) device-resident hot path: W warm-up steps, then exactly K timed steps
    run_chunks(max(args.warmup, 2) + (max(args.warmup, 2) % 2))        # even: the overlay's fwd/rev cycle stays aligned
    barrier()
    state.update(tracked=0, players=0, frames=0)
    ctx.launch_count(reset=True)
    if sampler:
        sampler.mark()
    if args.profile_region and not args.profile_4k:
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_chunks(args.steps)
    e1.record()
    barrier()
    if args.profile_region and not args.profile_4k:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    fps_dev = world * F * args.steps / (ms_dev / 1e3)
    per_frame = {"tracked_per_frame": state["tracked"] / max(state["frames"], 1),
                 "team_classified_per_frame": state["players"] / max(state["frames"], 1),
                 "candidates_per_frame": overlay.count(0)}
    if state["g"] % 2:
        run_chunks(1)

    # ---- (ii) end to end through the drop-in's public call: a clip of host frames in (pinned staging + H2D inside),
    # per-frame results out
    # The clip lives in page-locked host memory, frame by frame (what a decoder writing into buffers from hvb_host_alloc
    # leaves behind): Detector.upload copies such frames to the device from where they lie.  The same call on PAGEABLE
    # frames (staged through hvb_stage_frames into a pinned buffer first) is timed right after, as extra.e2e_from_pageable_frames_fps.
    frames_pinned = torch.from_numpy(frames).pin_memory()
    frames_p = frames_pinned.numpy()

    def run_e2e(n, src):
        g0 = state["g"]
        state["g"] += n
        clip = [src[i] for ids in pingpong(F, n, g0) for i in ids]
        return sum(1 for _ in vp.process_video_chunked(clip, chunk=F, initialize=False))

    def timed_e2e(src):
        run_e2e(2, src)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_out = run_e2e(args.steps, src)
        e1.record()
        barrier()
        assert n_out == F * args.steps
        ms = max_over_ranks(e0.elapsed_time(e1))
        if state["g"] % 2:
            run_e2e(1, src)
        return world * F * args.steps / (ms / 1e3)

    # three passes of K steps each, the median reported: on the shared GPU boxes one pass in three or four loses up to half
    # its rate to something outside the process (same library, same box, back to back: 714 / 1676 frames/s in run r02za) while
    # the device-timed figure does not move; every pass is listed in extra.e2e_passes_fps
    e2e_passes = [timed_e2e(frames_p) for _ in range(3)]
    fps_e2e = float(np.median(e2e_passes))
    fps_e2e_pageable = timed_e2e(frames)
    md = det.max_det
    n_team = int(round(per_frame["team_classified_per_frame"] * F))
    h2d = frames.nbytes + n_team * (16 + 4)
    d2h = F * md * (16 + 4 + 4) + F * 4 + F * md * 8 + F * 4 + n_team * 80

    # ---- roofline of the DOMINANT libhvb kernel of the step: the K5 conv epilogue (bias_act_kernel).  Every launch of
    # three more forwards is bracketed by CUDA events on its launching stream.
    peak, peak_src = measured_peaks()
    det.head_hook = None
    runner = det.runner
    k5 = None
    if runner is not None:
        runner.epi_log = []
        for _ in range(3):
            path.detect_device(fwd_dev)
        torch.cuda.synchronize()
        log, runner.epi_log = runner.epi_log, None
        k5_bytes = sum(b for b, _, _ in log)
        k5_ms = sum(a.elapsed_time(b) for _, a, b in log)
        k5 = {"launches_per_step": len(log) // 3, "bytes_per_step": k5_bytes // 3, "ms_per_step": k5_ms / 3,
              "achieved": k5_bytes / (k5_ms / 1e3) / 1e9}
    # K1a at the launch size of the step: back-to-back launches between one pair of events (inputs + outputs > L2)
    plan = det.plan(F, H, W, _ffi.LB_WHOLE)
    out = plan.run(fwd_dev)
    torch.cuda.synchronize()
    reps = max(args.steps, 20)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(reps):
        plan.run(fwd_dev, out)
    eb.record()
    torch.cuda.synchronize()
    k1_ms = ea.elapsed_time(eb) / reps
    k1_bytes = plan.read_bytes + plan.write_bytes
    del out

    # ---- per-stage device times of one chunk (CUDA events on the main stream, stages back to back, no overlap) — the GPU
    # column next to cpu_baseline.stage_ms_per_frame
    from hvb.tracker import DeviceByteTrack
    det.head_hook = overlay.to_device(dev, [list(range(F))])
    trk = DeviceByteTrack(0.25, 30, 0.8, 30, 2, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    acc = [0.0] * 5
    for it in range(4):
        det.head_hook.begin_chunk(F)
        ev[0].record()
        x = plan.class_views(plan.run(fwd_dev))[0]
        ev[1].record()
        heads = det.forward_heads(x)
        ev[2].record()
        meta_h, meta_d = det._meta_dev(plan, 0)
        xyxy, cf, cl, cnt, _ = det._decode(heads, meta_h, meta_d, F)
        ev[3].record()
        row, tid, tc = trk.update_chunk_device(xyxy, cf, cl, cnt, min_conf=0.4, class_mask=0b11)
        ev[4].record()
        path.team_device(fwd_dev, sk_dev, skf_dev)
        ev[5].record()
        torch.cuda.synchronize()
        if it:                                     # first pass is a warm-up
            for k in range(5):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) / 3 / F
    gpu_stage_ms = dict(zip(("letterbox+preprocess (K1a)", "yolo_forward (cuDNN convs + K5/K6, + planted scatter)",
                             "decode+nms+scale (K2a, %d candidates per frame)" % (PLAYERS * DUP),
                             "bytetrack (K7, %d frames in order)" % F,
                             "crops+team_predict (K3a/K3b, MobileNetV3, K4a; %d crops per frame)" % (PLAYERS - 1)),
                            (round(a, 5) for a in acc)))
    k7_us_per_frame = 1e3 * acc[3]
    del trk

    extra = {"fit": fit, "gpu_stage_ms_per_frame": gpu_stage_ms, "per_frame": per_frame, "tracker": args.tracker,
             "e2e_source": "frames in page-locked host memory, one H2D copy per chunk straight from them (no staging copy); median of three passes of K steps (extra.e2e_passes_fps)",
             "e2e_passes_fps": [round(v, 1) for v in e2e_passes], "e2e_from_pageable_frames_fps": fps_e2e_pageable, "staging_threads": det.staging_threads,
             "k7_bytetrack_us_per_frame_step": round(k7_us_per_frame, 2), "host_threads_per_rank": per_rank}

    # ---- the reference's actual loop shape: one frame at a time through process_frame (detect -> track -> crops -> predict)
    det.head_hook = hook_frames
    for tag, flag in (("eager", False), ("cuda_graph", True)):
        det.cuda_graph = flag
        vpf = VideoProcessor(device=dev, config=Config(), detector=det, team_classifier=path.classifier_router(), tracker=args.tracker)
        for _ in range(3):
            vpf.process_frame(frames[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            vpf.process_frame(frames[0])
        extra["frame_at_a_time_process_frame_fps_" + tag] = 20 / (time.perf_counter() - t0)
        if args.profile_dropin and flag:
            import cProfile, io, pstats
            pf = cProfile.Profile()
            pf.enable()
            for _ in range(20):
                vpf.process_frame(frames[0])
            pf.disable()
            buf = io.StringIO()
            pstats.Stats(pf, stream=buf).sort_stats("cumulative").print_stats(60)
            print("==== process_frame x20 (cuda graph)\n" + buf.getvalue(), file=sys.stderr)
        del vpf
    det.cuda_graph = False
    # the drop-in at its default chunk of 32 frames
    det.head_hook = overlay.to_device(dev, [list(range(32)), list(range(32, 64))] if F >= 64 else [list(range(F))])
    ch = 32 if F >= 64 else F
    clip = [frames[i] for k in range(12) for i in (range(32) if k % 2 == 0 else range(32, 64))] if F >= 64 else [frames[i] for k in range(12) for i in range(F)]
    vpc = VideoProcessor(device=dev, config=Config(), detector=det, team_classifier=path.classifier_router(), tracker=args.tracker)
    list(vpc.process_video_chunked(clip[:4 * ch], chunk=ch, initialize=False))
    barrier()
    runs = []
    prof = None
    if args.profile_dropin:                 # host-side profile of the drop-in's main thread (tools: where the 2x against e2e goes)
        import cProfile
        prof = cProfile.Profile()
    for _ in range(3):                      # a 384-frame clip lasts a quarter of a second: median of three passes
        t0 = time.perf_counter()
        if prof is not None:
            prof.enable()
        n_out = sum(1 for _ in vpc.process_video_chunked(clip, chunk=ch, initialize=False))
        if prof is not None:
            prof.disable()
        runs.append(n_out / (time.perf_counter() - t0))
    if prof is not None:
        import io, pstats
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(45)
        pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(30)
        print(buf.getvalue(), file=sys.stderr)
    extra["clip_chunked_drop_in_fps_chunk32"] = float(np.median(runs))
    extra["clip_chunked_drop_in_fps_chunk32_runs"] = [round(r, 1) for r in runs]
    del vpc
    det.head_hook = None
    if world > 1:                                   # the host-side figures of every rank
        for k in ("frame_at_a_time_process_frame_fps_cuda_graph", "clip_chunked_drop_in_fps_chunk32"):
            t = torch.tensor([extra[k]], device=dev)
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
            extra[k + "_ranks"] = {"min": float(lo.item()), "mean": float(t.item()) / world, "max": float(hi.item())}

    # ---- secondary workload: 4K sliced puck path (C4): planted pucks, cross-slice duplicates in the tile overlaps
    roofline_4k = None
    if args.with_4k:
        roofline_4k = bench_4k(args, rank, world, dev, barrier, max_over_ranks, peak, extra, path)
    # ---- BASELINE config 1: 1280x720, 60 frames, YOLOv8n (nc=1), sliced puck detection; CPU arm beside it at N=1
    if args.with_c1 and rank == 0:
        bench_c1(args, world, dev, extra)

    # ---- CPU baseline beside it (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores_all = os.cpu_count() or 1
        torch.set_num_threads(cores_all)
        nf = args.ref_frames
        ref = CpuReferencePath(overlay)
        ref.fit(frames[:nf + 2], boxes[:nf + 2])
        ref.step(frames[:2], [0, 1])                              # untimed: confirms the tracks (minimum_consecutive_frames = 2)
        ref.stage_s = {k: 0.0 for k in ref.STAGES}; ref.stage_frames = 0; ref.n_tracked = 0
        t0 = time.perf_counter()
        ref.step(frames[2:nf + 2], list(range(2, nf + 2)))
        dt = time.perf_counter() - t0
        stage_ms = ref.stage_ms_per_frame()
        tracked_cpu = ref.n_tracked / max(ref.stage_frames, 1)
        import cv2
        cv_threads = cv2.getNumThreads()
        torch.set_num_threads(1); cv2.setNumThreads(1)           # per-core normalisation
        t0 = time.perf_counter()
        ref.step(frames[nf + 2:nf + 3], [nf + 2])
        dt1 = time.perf_counter() - t0
        torch.set_num_threads(cores_all); cv2.setNumThreads(cv_threads)
        cpu = {"value": nf / dt, "unit": UNIT, "cores": cores_all, "kind": "port", "torch_threads": cores_all, "cv2_threads": cv_threads,
               "value_1_thread": 1.0 / dt1, "stage_ms_per_frame": stage_ms, "tracked_per_frame": tracked_cpu,
               "sample": "%d consecutive frames of the %d-frame clip, one frame at a time, same stages on the host: cv2 letterbox, "
                         "YOLOv8m CPU forward (torch, %d threads), the same planted candidates, restated decode + real torchvision "
                         "NMS, restated ByteTrack (scipy assignment), reference colour + MobileNetV3 features, predict + vote"
                         % (nf, F, cores_all)}

    k1_roof = {"kernel": "letterbox_kernel<false> (K1a, 1080p->736x1280, %d frames/launch)" % F, "bound": "hbm",
               "achieved": k1_bytes / (k1_ms / 1e3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
               "frac": k1_bytes / (k1_ms / 1e3) / 1e9 / peak, "traffic": ncu_traffic(["K1a_1080p_x%d_%s" % (F, r) for r in ("r02zf", "r02f")]),
               "algorithmic_bytes_per_launch": int(k1_bytes), "avg_launch_ms": k1_ms}
    if k5 is not None:
        n5 = k5["launches_per_step"]
        t5 = ncu_traffic(["K5_bias_act_step_x%d_%s" % (F, r) for r in ("r02zf", "r02f")], launches=n5)   # summed over the launches of one step
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
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "players_per_frame": PLAYERS,
                       "planted_candidates_per_frame": PLAYERS * DUP, "team_boxes": "detected",
                       "backbones": "convolutions in torch/cuDNN (fp32 storage, conv+bn folded, channels_last, cudnn.benchmark, TF32 for YOLO / "
                                    "TF32 off for MobileNetV3); everything between the YOLO convolutions (bias, SiLU, residual, concat, "
                                    "upsample, layer 0) in libhvb K5 kernels; six pointwise convolutions as single K6 launches (tcgen05 TF32 GEMM + epilogue)",
                       "l2": "inputs larger than L2 (%.0f MB frames + %.0f MB letterboxed per step)" % (frames.nbytes / 1e6, k1_bytes / 1e6),
                       "parallelism": "one clip per GPU, frame chunks in order, no data-path collective; one NCCL feature all-gather at fit"},
            "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_k1a": k1_roof,          # SURVEY 8(d)'s own memory-bound kernel of this workload, beside the dominant one
            "roofline_4k": roofline_4k,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line))


def bench_4k(args, rank, world, dev, barrier, max_over_ranks, peak, extra, path):
    """C4: 4K frames -> K1b slice letterbox (40 tiles, 5 shape classes) -> YOLOv8n per class -> planted pucks -> K2a ->
    gather -> K2b cross-slice merge.  Returns the top-level roofline block of the 4K path (K1b + K2a at 640 tiles)."""
    import torch
    from hvb import _ffi
    from hvb.pipeline import SlicedPuckPath
    from hvb.synth import PlantedOverlay
    F4 = args.chunk_4k
    f4, boxes4, _, pucks = synth_clip(7 + rank, F4, 2160, 3840, 2.0)
    # per frame: the puck + two more small objects near tile seams, so that overlapping tiles see the same object
    rng = np.random.default_rng(5 + rank)
    objs = []
    for p in pucks:
        extra_xy = np.stack([rng.choice([512, 1024, 1536, 2048, 2560, 3072], 2) + rng.uniform(20, 100, 2),
                             rng.choice([512, 1024, 1536], 2) + rng.uniform(20, 100, 2)], 1)
        eb = np.hstack([extra_xy, extra_xy + rng.uniform(14, 24, (2, 2))])
        objs.append(np.vstack([p[None], eb]))
    ov4 = PlantedOverlay.sliced(9 + rank, (2160, 3840), 640, objs, [np.zeros(3, int)] * F4, nc=1, dup=2)
    f4_dev = torch.from_numpy(f4).to(dev)
    puck = SlicedPuckPath(dev, "n", 1, 0.4)
    plan4 = puck.detector.plan(F4, 2160, 3840, _ffi.LB_SLICE_EXACT, 640, (640, 640), (128, 128))
    puck.detector.head_hook = ov4.to_device(dev, [list(range(F4))], plan=plan4)
    k4 = max(2, args.steps // 2)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            torch.cuda.profiler.start()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if profile:
            torch.cuda.profiler.stop()
        return max_over_ranks(e0.elapsed_time(e1))

    ms4_eager = timed(lambda: puck.process_chunk_device(f4_dev), k4, 2, profile=args.profile_region and args.profile_4k)
    # the sliced path is launch-bound (~900 launches per chunk over 5 tile shape classes): replay it as one CUDA graph
    ms4 = timed(lambda: puck.process_chunk_device(f4_dev, graph=True), k4, 2)
    res = puck.process_chunk_device(f4_dev, sync=True)
    keep, seg = res[3].cpu().numpy(), res[4].cpu().numpy()
    merged_per_frame = float(seg[-1]) / F4
    kept_per_frame = float((keep == 1).sum()) / F4
    # end to end through the public sliced API: pinned 4K frames in, per-frame Detections out
    pinned4 = torch.from_numpy(f4).pin_memory()

    def run_e2e4(k):
        return sum(len(r) for r in puck.process_stream(pinned4 for _ in range(k)))

    run_e2e4(4)
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    run_e2e4(k4)
    eb.record()
    barrier()
    ms4_e2e = max_over_ranks(ea.elapsed_time(eb))
    # K1b and K2a at the 4K launch sizes, back to back between one pair of events
    o4 = plan4.run(f4_dev)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(20):
        plan4.run(f4_dev, o4)
    eb.record()
    torch.cuda.synchronize()
    k1b_ms = ea.elapsed_time(eb) / 20
    k1b_bytes = plan4.read_bytes + plan4.write_bytes
    # per-stage device times of one eager chunk (the 5 shape classes' forwards and K2a launches summed)
    stage4 = {}
    for it in range(3):
        puck.slicer.stage_events = marks = []
        puck.process_chunk_device(f4_dev)
        torch.cuda.synchronize()
        puck.slicer.stage_events = None
        if it:
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                stage4[name] = stage4.get(name, 0.0) + a.elapsed_time(b) / 2
    tiles = F4 * int(plan4.tiles_per_frame)
    extra.update({"c4_4k_sliced_frames_per_sec": world * F4 * k4 / (ms4 / 1e3),
                  "c4_4k_sliced_e2e_frames_per_sec": world * F4 * k4 / (ms4_e2e / 1e3),
                  "c4_4k_sliced_frames_per_sec_eager_launches": world * F4 * k4 / (ms4_eager / 1e3),
                  "c4_e2e_h2d_bytes_per_step": int(f4.nbytes), "c4_mode": "whole chunk replayed as one CUDA graph",
                  "c4_frames_per_step": F4, "c4_tiles_per_frame": int(plan4.tiles_per_frame),
                  "c4_merged_detections_per_frame_before_cross_slice_nms": merged_per_frame,
                  "c4_detections_per_frame_after_cross_slice_nms": kept_per_frame})
    puck.detector.head_hook = None
    # ---- BASELINE config 5, one GPU's share: a 4K clip through the FULL pipeline — players (whole-frame letterbox to
    # 720x1280, YOLOv8m, planted candidates, K2a, K7, team stage on the tracked skaters) AND the sliced puck path on the
    # same device-resident chunk, per step.  (The global fit's all-gather is `extra.fit`.)
    from hvb.video import Config, VideoProcessor
    det = path.detector
    cls4 = [np.array([0] * (PLAYERS - 1) + [1])] * F4
    ov_players = PlantedOverlay.whole_frame(31 + rank, (2160, 3840), IMGSZ, boxes4, cls4, nc=2, dup=DUP)
    hook_players = ov_players.to_device(dev, [list(range(F4))])
    hook_pucks = ov4.to_device(dev, [list(range(F4))], plan=plan4)
    vp4 = VideoProcessor(device=dev, config=Config(), detector=det, team_classifier=path.classifier_router(), tracker=args.tracker)
    tracked = [0, 0]

    def full_step():
        det.head_hook = hook_players
        for r in vp4.process_chunks([f4_dev]):
            tracked[0] += len(r.detections); tracked[1] += 1
        det.head_hook = None
        puck.detector.head_hook = hook_pucks
        out = puck.process_chunk_device(f4_dev, graph=True)
        puck.detector.head_hook = None
        return out

    ms5 = timed(full_step, k4, 3)
    extra["c5_4k_full_pipeline_frames_per_sec"] = world * F4 * k4 / (ms5 / 1e3)
    extra["c5_note"] = ("per step and GPU: %d 4K frames, players (K1a 4K->720x1280, YOLOv8m, K2a, K7, team stage; %.1f tracked per frame) + sliced "
                        "puck path (640 tiles, one graph replay), device-resident" % (F4, tracked[0] / max(tracked[1], 1)))
    return {"workload": "C4: 4K sliced puck detection, %d frames = %d tiles per step, 3 planted objects per frame (x2 candidates, cross-slice duplicates)"
                        % (F4, F4 * int(plan4.tiles_per_frame)),
            "kernel": "letterbox_kernel<true> (K1b slice letterbox, exact 5-shape-class mode)", "bound": "hbm",
            "achieved": k1b_bytes / (k1b_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": k1b_bytes / (k1b_ms / 1e3) / 1e9 / peak,
            "algorithmic_bytes_per_launch": int(k1b_bytes), "avg_launch_ms": k1b_ms, "traffic": ncu_traffic(["K1b_4k_x%d_r02zf" % F4, "K1b_4k_x%d" % F4]),
            "frames_per_sec": world * F4 * k4 / (ms4 / 1e3), "e2e_frames_per_sec": world * F4 * k4 / (ms4_e2e / 1e3),
            "stage_ms_per_chunk": {k: round(v, 4) for k, v in stage4.items()},
            "k2a": {"launches_per_chunk": len(plan4.classes), "tiles": tiles, "candidates_per_frame": merged_per_frame,
                    "us_per_chunk": round(1e3 * stage4.get("K2a decode + NMS", 0.0), 1),
                    "note": "latency-bound (a few MB of class logits per launch): reported as time, not as a fraction"}}


def bench_c1(args, world, dev, extra):
    """BASELINE config 1: PUCK_DETECTION on a synthetic 1280x720 60-frame clip, random-init YOLOv8n (nc=1), sliced
    (6 tiles, 0.2 overlap) — the GPU path on the whole clip, and (N=1) the restated CPU slicer beside it on 2 frames."""
    import torch
    from hvb import _ffi
    from hvb.pipeline import SlicedPuckPath
    from hvb.synth import PlantedOverlay
    n = 60
    f1, _, _, pucks = synth_clip(3, n, 720, 1280, 0.67)
    objs = [p[None] for p in pucks]
    ov = PlantedOverlay.sliced(4, (720, 1280), 640, objs, [np.zeros(1, int)] * n, nc=1, dup=2)
    puck = SlicedPuckPath(dev, "n", 1, 0.4)
    plan = puck.detector.plan(n, 720, 1280, _ffi.LB_SLICE_EXACT, 640, (640, 640), (128, 128))
    puck.detector.head_hook = ov.to_device(dev, [list(range(n))], plan=plan)
    pinned = torch.from_numpy(f1).pin_memory()
    list(puck.process_stream(pinned for _ in range(4)))     # both frame buffers, all three pinned result sets, the graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 4
    dets = [d for out in puck.process_stream(pinned for _ in range(k)) for d in out]
    dt = time.perf_counter() - t0
    c1 = {"workload": "C1: 1280x720, 60-frame clip, YOLOv8n nc=1, sliced (6 tiles per frame), 1 planted puck per frame",
          "gpu_e2e_frames_per_sec": k * n / dt, "detections_per_frame": sum(len(d) for d in dets) / len(dets)}
    puck.detector.head_hook = None
    if world == 1 and not args.no_cpu_baseline:
        from hvb.models import build_yolov8
        from hvb.models.yolov8 import fuse_conv_bn
        from oracle import supervision_restated as svr, ultralytics_restated as ur
        yolo = fuse_conv_bn(build_yolov8("n", 1, 0)).eval()
        offs = svr.generate_offsets((1280, 720), (640, 640), (0.2, 0.2))

        def cpu_frame(fid):
            parts = []
            for t, off in enumerate(offs):
                tile = np.ascontiguousarray(svr.crop_image(f1[fid], off))
                x = torch.from_numpy(ur.preprocess([ur.letterbox(tile, 640, auto=True)]))
                with torch.no_grad():
                    heads = [h.clone() for h in yolo(x)]
                ov.apply_host(heads, fid, t)
                xyxy, conf, cls = ur.predict_from_head(heads, 1, tuple(x.shape[2:]), [tile.shape[:2]], 0.4)[0]
                parts.append((svr.move_boxes(xyxy, off[:2]), conf, cls))
            xy = np.concatenate([p[0] for p in parts]); cf = np.concatenate([p[1] for p in parts]); cl = np.concatenate([p[2] for p in parts])
            return xy[svr.with_nms(xy, cf, cl, 0.1)] if len(xy) else xy

        cpu_frame(0)
        t0 = time.perf_counter()
        nd = sum(len(cpu_frame(i)) for i in (0, 1))
        c1["cpu_frames_per_sec"] = 2 / (time.perf_counter() - t0)
        c1["cpu_detections_per_frame"] = nd / 2
        c1["cpu_cores"] = os.cpu_count()
    extra["c1_720p_yolov8n_sliced"] = c1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hvb", choices=["hvb", "reference"])
    ap.add_argument("--chunk", type=int, default=64, help="1080p frames per step per GPU")
    ap.add_argument("--chunk-4k", type=int, default=16, help="4K frames per step of the sliced puck path (640 tiles per step)")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per CPU-reference step (bounded sample)")
    ap.add_argument("--tracker", default="device", choices=["device", "host"], help="K7 (device ByteTrack) or the host tracker")
    ap.add_argument("--no-4k", dest="with_4k", action="store_false")
    ap.add_argument("--no-c1", dest="with_c1", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-4k", action="store_true", help="with --profile-region: bracket the eager 4K sliced steps instead of the 1080p steps")
    ap.add_argument("--profile-dropin", action="store_true", help="cProfile of the main thread over the timed process_video_chunked passes (to stderr)")
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
