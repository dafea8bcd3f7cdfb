"""Per-kernel timings of libhvb at the BASELINE config sizes (CUDA events around back-to-back
launches on torch's current stream).  Prints one JSON line per kernel; used for profiles/*.md.

    python tools/kernel_bench.py [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))
from hvb import _ffi  # noqa: E402
from hvb.runtime import get_context  # noqa: E402
from hvb.synth import planted_head, random_boxes, rink_frame  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps):
    if reps <= 0:                       # --profile: exactly one launch of everything (ncu captures), no timing
        fn(); torch.cuda.synchronize()
        return 1.0
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, ms, nbytes=None, flops=None, **kw):
    d = {"kernel": name, "us": round(1e3 * ms, 2)}
    if nbytes:
        d["algorithmic_MB"] = round(nbytes / 1e6, 3)
        d["GB_per_s"] = round(nbytes / (ms / 1e3) / 1e9, 1)
        d["frac_hbm_peak"] = round(d["GB_per_s"] / PEAK, 3)
    if flops:
        d["TFLOP_per_s"] = round(flops / (ms / 1e3) / 1e12, 2)
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="comma-separated sections: k1,k2,k3,k4,k4b,track,k5,k6,k8")
    ap.add_argument("--profile", action="store_true", help="one launch per kernel configuration, no timing (for ncu)")
    args = ap.parse_args()
    if args.profile:
        args.reps = 0
    only = set(args.only.split(",")) if args.only else None

    def want(tag):
        return only is None or tag in only
    ctx = get_context(0)
    rng = np.random.default_rng(0)
    R = args.reps

    def sec_k1():
        # K1a / K1b
        for tag, n, h, w, mode, imgsz in (("K1a 1080p->736x1280 x32 (bench.py's launch)", 32, 1080, 1920, _ffi.LB_WHOLE, 1280),
                                          ("K1a 1080p->736x1280 x16", 16, 1080, 1920, _ffi.LB_WHOLE, 1280),
                                          ("K1a 1080p->736x1280 x64", 64, 1080, 1920, _ffi.LB_WHOLE, 1280),
                                          ("K1a 720p->384x640 x64 (2x area path)", 64, 720, 1280, _ffi.LB_WHOLE, 640),
                                          ("K1b 4K sliced exact x4", 4, 2160, 3840, _ffi.LB_SLICE_EXACT, 640),
                                          ("K1b 4K sliced exact x16 (bench.py's 4K launch)", 16, 2160, 3840, _ffi.LB_SLICE_EXACT, 640),
                                          ("K1b 4K sliced uniform x8", 8, 2160, 3840, _ffi.LB_SLICE_UNIFORM, 640)):
            frames = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
            plan = ctx.letterbox_plan(n, h, w, mode, imgsz)
            out = plan.run(frames)
            report(tag, timed(lambda: plan.run(frames, out), R), plan.read_bytes + plan.write_bytes)
            del frames, out

    def sec_k2():
        # K2a: C2 (1080p, nc=2, 16 images), C4-like (640x640 tiles, nc=1, 160 images)
        for tag, B, hw, nc, n_gt in (("K2a decode+NMS C2 736x1280 nc=2 x16", 16, (736, 1280), 2, 12),
                                     ("K2a decode+NMS tiles 640x640 nc=1 x160", 160, (640, 640), 1, 3)):
            lv = [(hw[0] // s, hw[1] // s) for s in (8, 16, 32)]
            one = []
            for _ in range(4):
                cx, cy = rng.uniform(60, hw[1] - 60, n_gt), rng.uniform(60, hw[0] - 60, n_gt)
                bw, bh = rng.uniform(20, 100, n_gt), rng.uniform(30, 200, n_gt)
                gt = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
                one.append(planted_head(rng, lv, nc, gt, rng.integers(0, nc, n_gt)))
            levels = [torch.from_numpy(np.stack([one[i % 4][l] for i in range(B)])).cuda() for l in range(3)]
            meta = np.zeros((B,), _ffi.IMG_META)
            meta["gain"], meta["clip_w"], meta["clip_h"], meta["out_slot"] = 1.0, hw[1], hw[0], np.arange(B)
            meta_d = ctx.struct_to_device(meta)
            outs = ctx.decode_nms(levels, nc, 0.4, 0.7, 300, False, meta=meta)
            A = sum(a * b for a, b in lv)
            report(tag, timed(lambda: ctx.decode_nms(levels, nc, 0.4, 0.7, 300, False, meta=meta_d, out=outs, check_overflow=False), R),
                   B * A * nc * 4, full_head_MB=round(B * A * (64 + nc) * 4 / 1e6, 2), kept=int(outs[3].sum().item()))
        # K2b: 64 frames x ~120 merged detections
        n_seg, per = 64, 120
        boxes = np.concatenate([random_boxes(rng, per, 3840, 2160, 10, 40, np.float64)[0] for _ in range(n_seg)])
        conf = rng.permutation(np.linspace(0.3, 0.99, n_seg * per)).astype(np.float32)
        cls = np.zeros(n_seg * per, np.int32)
        seg = (np.arange(n_seg + 1) * per).astype(np.int32)
        bd, cd_, kd, sd = [torch.from_numpy(x).cuda() for x in (boxes, conf, cls, seg)]
        report("K2b merge NMS 64 frames x 120 dets", timed(lambda: ctx.merge_nms(bd, cd_, kd, sd, n_seg, n_seg * per, 0.1), R),
               n_seg * per * 40)

    def sec_k3():
        # K3a / K3b: 12 crops per 1080p frame, 64 frames
        nf = 64
        frs, bxs, fidx = [], [], []
        for i in range(nf):
            f, b, _, _ = rink_frame(rng, 1080, 1920, 12)
            frs.append(f); bxs.append(b); fidx.append(np.full(len(b), i, np.int32))
        frames = torch.from_numpy(np.stack(frs)).cuda()
        bx = torch.from_numpy(np.concatenate(bxs).astype(np.float32)).cuda()
        fi = torch.from_numpy(np.concatenate(fidx)).cuda()
        m = bx.shape[0]
        cd = ctx.crops_from_boxes(bx, fi, 1080, 1920)
        desc = cd.cpu().numpy().view(_ffi.CROP_DESC)[:m]
        roi_px = sum((int(h * 0.6) - int(h * 0.1)) * (int(w * 0.8) - int(w * 0.2)) if (h >= 40 and w >= 20) else h * w
                     for h, w in zip(desc["h"], desc["w"]))
        feat = ctx.empty((m, 49), torch.float64)
        report("K3a colour features %d crops" % m, timed(lambda: ctx.color_features(frames, cd, m, out_feat=feat), R),
               roi_px * 3 + m * 392, roi_kpx_per_crop=round(roi_px / m / 1e3, 2))
        report("K3b MobileNetV3 prep %d crops" % m, timed(lambda: ctx.mnv3_preprocess(frames, cd, m), R), roi_px * 3 + m * 98304)
        seg_px = sum((int(h * 0.6) - int(h * 0.2)) * (int(w * 0.7) - int(w * 0.3)) for h, w in zip(desc["h"], desc["w"]))
        report("K3c jersey colour stats (segmentation rectangle) %d crops" % m,
               timed(lambda: ctx.jersey_color_stats(frames, cd, m, _ffi.ROI_SEGMENT), R), seg_px * 3 + m * 120,
               roi_kpx_per_crop=round(seg_px / m / 1e3, 2))
        report("crops_from_boxes %d" % m, timed(lambda: ctx.crops_from_boxes(bx, fi, 1080, 1920), R))

    def sec_k4():
        # K4a: standardise + affinity, tensor-core Gram
        for n in (250, 2000, 8192):
            x = torch.from_numpy(rng.normal(0, 1, (n, 625))).cuda()
            mean, scale, xs = ctx.standardize(x)
            report("K4a standardize N=%d" % n, timed(lambda: ctx.standardize(x), R), n * 625 * 8 * 3)
            report("K4a gram tcgen05 N=%d (3xTF32, K'=1920)" % n, timed(lambda: ctx.gram_tc(xs), R), flops=2.0 * n * n * 1920,
                   useful_TFLOP_per_s=None)
            if n <= 2000:
                report("K4a affinity mode0 (tcgen05+refine) N=%d" % n, timed(lambda: ctx.gram_affinity(xs, 1.0, 0), max(R // 4, 2)), flops=2.0 * n * n * 625)
                report("K4a affinity mode1 (fp64) N=%d" % n, timed(lambda: ctx.gram_affinity(xs, 1.0, 1), max(R // 4, 2)), flops=3.0 * n * n * 625)

    def sec_k4b():
        # K4b: 8 clips x (40 tracks x 40 detections)
        T = D = 40
        P = 8
        a = torch.from_numpy(np.concatenate([random_boxes(rng, T, 1920, 1080, 40, 200, np.float64)[0] for _ in range(P)])).cuda()
        b = torch.from_numpy(np.concatenate([random_boxes(rng, D, 1920, 1080, 40, 200, np.float64)[0] for _ in range(P)])).cuda()
        ao = torch.from_numpy((np.arange(P + 1) * T).astype(np.int32)).cuda()
        bo = torch.from_numpy((np.arange(P + 1) * D).astype(np.int32)).cuda()
        oo = torch.from_numpy((np.arange(P) * T * D).astype(np.int64)).cuda()
        report("K4b IoU cost 8 clips x 40x40", timed(lambda: ctx.iou_cost(a, b, None, ao, bo, oo, P, T, D, P * T * D, 2), R), P * (T + D) * 32 + P * T * D * 8)

    def sec_track():
        # ByteTrack association for 8 clips x ~40 detections per frame: independent trackers (one K4b launch + copy per
        # clip and round) vs MultiClipByteTrack (one batched launch per round for all clips)
        import time
        from hvb.detections import Detections
        from hvb.tracker import ByteTrack, MultiClipByteTrack
        kw = dict(track_activation_threshold=0.25, lost_track_buffer=30, minimum_matching_threshold=0.8, frame_rate=30,
                  minimum_consecutive_frames=2)
        n_clips, n_frames, n_obj = 8, 60, 40
        r = np.random.default_rng(9)
        clips = []
        for c in range(n_clips):
            cx, cy = r.uniform(200, 1700, n_obj), r.uniform(200, 900, n_obj)
            w, h = r.uniform(40, 110, n_obj), r.uniform(100, 250, n_obj)
            fr = []
            for f in range(n_frames):
                cx += r.uniform(-6, 6, n_obj); cy += r.uniform(-6, 6, n_obj)
                b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
                fr.append((b, r.uniform(0.45, 0.95, n_obj).astype(np.float32)))
            clips.append(fr)
        mk = lambda c, f: Detections(xyxy=clips[c][f][0].copy(), confidence=clips[c][f][1].copy(), class_id=np.zeros(n_obj, int))
        singles = [ByteTrack(**kw) for _ in range(n_clips)]
        t0 = time.perf_counter()
        for f in range(n_frames):
            for c in range(n_clips):
                singles[c].update_with_detections(mk(c, f))
        t_single = (time.perf_counter() - t0) / n_frames
        multi = MultiClipByteTrack(n_clips, **kw)
        t0 = time.perf_counter()
        for f in range(n_frames):
            multi.update_with_detections([mk(c, f) for c in range(n_clips)])
        t_multi = (time.perf_counter() - t0) / n_frames
        print(json.dumps({"kernel": "ByteTrack 8 clips x 40 detections, host logic + K4b costs", "ms_per_frame_independent": round(1e3 * t_single, 3),
                          "ms_per_frame_lockstep_batched_k4b": round(1e3 * t_multi, 3),
                          "clip_frames_per_s_lockstep": round(n_clips / t_multi, 1)}), flush=True)
        # K7: the same clips through the device tracker — chunks of 32 frames of all 8 clips per launch, and (second line)
        # the drop-in's shape: ONE clip, 12 detections per frame, 64-frame chunks
        from hvb.tracker import DeviceByteTrack
        for tag, nc_, nobj, ch in (("8 clips x 40 detections, 30-frame chunks", n_clips, n_obj, 30), ("1 clip x 12 detections, 60-frame chunks", 1, 12, 60)):
            trk = DeviceByteTrack(n_clips=nc_, **kw)
            md = 64
            packs = []
            for f0 in range(0, n_frames, ch):
                xy = np.zeros((nc_ * ch, md, 4), np.float32); cf = np.zeros((nc_ * ch, md), np.float32); cnt = np.zeros(nc_ * ch, np.int32)
                for c in range(nc_):
                    for f in range(ch):
                        xy[c * ch + f, :nobj], cf[c * ch + f, :nobj], cnt[c * ch + f] = clips[c][f0 + f][0][:nobj], clips[c][f0 + f][1][:nobj], nobj
                packs.append([torch.from_numpy(a).cuda() for a in (xy, cf, cnt)])
            def run():
                for xy, cf, cnt in packs:
                    trk.update_chunk_device(xy, cf, None, cnt)
            ms = timed(run, max(R // 4, 1) if R > 0 else 0)
            print(json.dumps({"kernel": "K7 device ByteTrack, " + tag, "us_per_launch": round(1e3 * ms / len(packs), 1),
                              "us_per_frame_step": round(1e3 * ms / n_frames, 2),
                              "clip_frames_per_s": round(nc_ * n_frames / (ms / 1e3), 1)}), flush=True)

    def sec_k5():
        # K5 backbone glue at the YOLOv8m / 1080p (736x1280 input) layer sizes, 32 frames per launch
        CL = torch.channels_last
        nb = 32
        shapes = (("layer0 out 48ch 368x640", 48, 368, 640), ("layer1 out 96ch 184x320", 96, 184, 320),
                  ("bottleneck 48ch 184x320", 48, 184, 320), ("p3 192ch 92x160", 192, 92, 160))
        for tag, c, h, w in shapes[:1] if args.profile else shapes:
            x = torch.randn(nb, c, h, w, device="cuda").contiguous(memory_format=CL)
            b = torch.randn(c, device="cuda")
            r = torch.randn(nb, c, h, w, device="cuda").contiguous(memory_format=CL)
            nbytes = x.numel() * 4
            report("K5 bias+SiLU in place x%d %s" % (nb, tag), timed(lambda: ctx.bias_act(x, b, "silu"), R), 2 * nbytes)
            report("K5 bias+SiLU+residual x%d %s" % (nb, tag), timed(lambda: ctx.bias_act(x, b, "silu", residual=r), R), 3 * nbytes)
            cat = torch.empty(nb, 2 * c, h, w, device="cuda").contiguous(memory_format=CL)
            report("K5 bias+SiLU -> dense + concat slice x%d %s" % (nb, tag),
                   timed(lambda: ctx.bias_act(x, b, "silu", out1=x, out2=cat, out2_off=c, c2_begin=0, c2_count=c), R), 3 * nbytes)
            del x, r, cat
        p5 = torch.randn(nb, 576, 23, 40, device="cuda").contiguous(memory_format=CL)
        p4 = torch.randn(nb, 384, 46, 80, device="cuda").contiguous(memory_format=CL)
        out = ctx.concat_nhwc([p5, p4], [1, 0])
        report("K5 upsample2x+concat x%d (576@23x40, 384@46x80)" % nb, timed(lambda: ctx.concat_nhwc([p5, p4], [1, 0], out=out), R),
               p5.numel() * 4 + p4.numel() * 4 + out.numel() * 4)
        y0 = torch.randn(nb, 288, 23, 40, device="cuda").contiguous(memory_format=CL)
        cat4 = ctx.sppf_pool_concat(y0)
        report("K5 SPPF 3x maxpool5 + concat x%d 288ch 23x40" % nb, timed(lambda: ctx.sppf_pool_concat(y0, out=cat4), R),
               y0.numel() * 4 * 5)
        xin = torch.rand(nb, 3, 736, 1280, device="cuda")
        wst = (np.random.default_rng(0).standard_normal((48, 3, 3, 3)) * 0.2).astype(np.float32)
        bst = np.zeros((48,), np.float32)
        ms = timed(lambda: ctx.stem_conv(xin, wst, bst), R)
        report("K5 stem conv 3->48 s2 + SiLU x%d 736x1280" % nb, ms, xin.numel() * 4 + nb * 48 * 368 * 640 * 4,
               flops=2 * 27 * 48 * nb * 368 * 640)
        del xin
        nt = 112                                                   # 4 x 4K frames: the 640x640 tile class of the sliced path
        xt = torch.rand(nt, 3, 640, 640, device="cuda")
        w16 = (np.random.default_rng(1).standard_normal((16, 3, 3, 3)) * 0.2).astype(np.float32)
        b16 = np.zeros((16,), np.float32)
        ms = timed(lambda: ctx.stem_conv(xt, w16, b16), R)
        report("K5 stem conv 3->16 s2 + SiLU x%d tiles 640x640 (YOLOv8n)" % nt, ms, xt.numel() * 4 + nt * 16 * 320 * 320 * 4,
               flops=2 * 27 * 16 * nt * 320 * 320)

    if want('k1'):
        sec_k1()
    if want('k2'):
        sec_k2()
    def sec_k6():
        # K6 fused pointwise conv vs what it replaces (cuDNN TF32 conv2d + K5 epilogue), YOLOv8m layers at 32 x 1080p
        CL = torch.channels_last
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        for tag, hw, cin, cout in (("L2.cv1 96->96 @184x320", (184, 320), 96, 96), ("L2.cv2 192->96 @184x320", (184, 320), 192, 96),
                                   ("L4.cv1 192->192 @92x160", (92, 160), 192, 192), ("L4.cv2 576->192 @92x160", (92, 160), 576, 192),
                                   ("L6.cv2 1152->384 @46x80", (46, 80), 1152, 384), ("L8.cv2 1152->576 @23x40", (23, 40), 1152, 576),
                                   ("Detect 64->64 @92x160", (92, 160), 64, 64)):
            x = torch.randn((32, cin, hw[0], hw[1]), device="cuda").contiguous(memory_format=CL)
            w = (torch.randn((cout, cin, 1, 1), device="cuda") / cin ** 0.5).contiguous(memory_format=CL)
            b = torch.randn(cout, device="cuda")
            out = torch.empty((32, cout, hw[0], hw[1]), device="cuda").contiguous(memory_format=CL)
            npix = 32 * hw[0] * hw[1]
            alg = npix * (cin + cout) * 4
            report("K6 fused " + tag, timed(lambda: ctx.pointwise_conv(x, w, b, "silu_fast", out1=out), R), alg)
            report("   cuDNN conv2d + K5 " + tag, timed(lambda: ctx.bias_act(torch.conv2d(x, w), b, "silu_fast"), R), alg)
            del x, out

    def sec_k8():
        # K8: the subspace-iteration pass over the normalised affinity, and the whole device spectral fit
        from hvb.spectral import DeviceSpectralClustering, _DeviceOps, spectral_embedding_subspace
        for n in (2000, 8192):
            g = torch.Generator(device="cuda").manual_seed(0)
            c = torch.randn((2, 32), device="cuda", dtype=torch.float64, generator=g) * 1.5
            x = torch.randn((n, 32), device="cuda", dtype=torch.float64, generator=g) + c[torch.arange(n, device="cuda") % 2]
            a = torch.exp(-torch.cdist(x, x) ** 2 / 32)
            m, dd = ctx.laplacian_normalize(a)
            report("K8 laplacian_normalize N=%d" % n, timed(lambda: ctx.laplacian_normalize(a), R), 2 * n * n * 8 + n * n * 8)
            v = torch.randn((8, n), device="cuda", dtype=torch.float64, generator=g)
            y = torch.empty_like(v)
            report("K8 sym_block_matvec N=%d (8 vectors)" % n, timed(lambda: ctx.sym_block_matvec(m, v, 0.1, y), R), n * n * 8 + 2 * 8 * n * 8,
                   flops=2 * 8 * n * n)
            if R > 0:
                info = {}
                torch.cuda.synchronize()
                import time as _t
                t0 = _t.perf_counter()
                spectral_embedding_subspace(a, 2, _DeviceOps(ctx), info=info)
                t1 = _t.perf_counter()
                lab = DeviceSpectralClustering(2, 10, 42).fit_predict(a)
                t2 = _t.perf_counter()
                torch.linalg.eigh(m)
                torch.cuda.synchronize()
                t3 = _t.perf_counter()
                print(json.dumps({"kernel": "K8 device spectral fit N=%d" % n, "embedding_ms": round(1e3 * (t1 - t0), 2),
                                  "embedding+kmeans_ms": round(1e3 * (t2 - t1), 2), "cusolver_eigh_ms": round(1e3 * (t3 - t2), 2),
                                  "outer_rounds": info["outer_iterations"], "matvecs": info["matvecs"],
                                  "residual": float(info["residuals"].max()), "labels_split": [int((lab == 0).sum()), int((lab == 1).sum())]}), flush=True)
            del a, m, x

    if want('k3'):
        sec_k3()
    if want('k8'):
        sec_k8()
    if want('k6'):
        sec_k6()
    if want('k4'):
        sec_k4()
    if want('k4b'):
        sec_k4b()
    if want('track'):
        sec_track()
    if want('k5'):
        sec_k5()


if __name__ == "__main__":
    main()
