"""Time the YOLOv8m forward of one bench step (32 x 736 x 1280, fp32, TF32 convolutions) with the pointwise layers on
K6 and on cuDNN + K5.  CUDA events on the launching stream, inputs > L2.

    python tools/forward_bench.py [--frames 32] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--per-layer", action="store_true")
    ap.add_argument("--skip-total", action="store_true")
    args = ap.parse_args()
    from hvb import get_context
    from hvb.models import build_yolov8
    from hvb.models.fused import FusedYOLOv8
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    ctx = get_context(0)
    model = build_yolov8("m", 2, 0)
    x = torch.rand(args.frames, 3, 736, 1280, device="cuda")
    out = {}
    for tag, pw in (() if args.skip_total else (("cudnn+k5", False), ("k6", True), ("cudnn+k5 again", False), ("k6 again", True))):
        run = FusedYOLOv8(model, ctx, pointwise_kernel=pw)
        for _ in range(3):
            run(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            run(x)
        b.record(); torch.cuda.synchronize()
        out[tag] = round(a.elapsed_time(b) / args.reps, 3)
    print(json.dumps({"yolov8m_forward_ms_per_%d_frames" % args.frames: out}))
    if args.per_layer:
        per_layer(ctx, model, x, args.reps)


def per_layer(ctx, model, x, reps):
    """Per-layer time of the pointwise layers INSIDE the forward, both ways (events wrapped around the product's own
    calls from here; the product code is untouched)."""
    from hvb.models import fused
    E = lambda: torch.cuda.Event(enable_timing=True)
    log, pending = [], []
    orig_pw, orig_raw, orig_epi = fused.FusedYOLOv8._pw, fused._Conv.raw, fused.FusedYOLOv8._epi

    def pw(self, k, x_, *a, **kw):
        e0, e1 = E(), E()
        e0.record()
        out = orig_pw(self, k, x_, *a, **kw)
        e1.record()
        log.append(("k6", k.cin, k.cout, x_.shape[2], e0, e1))
        return out

    def raw(self, x_):
        if self.pointwise:
            e0 = E()
            e0.record()
            pending.append((self, x_.shape[2], e0))
        return orig_raw(self, x_)

    def epi(self, x_, bias, *a, **kw):
        out = orig_epi(self, x_, bias, *a, **kw)
        if pending:
            k, h, e0 = pending.pop()
            e1 = E()
            e1.record()
            log.append(("cudnn+k5", k.cin, k.cout, h, e0, e1))
        return out

    fused.FusedYOLOv8._pw, fused._Conv.raw, fused.FusedYOLOv8._epi = pw, raw, epi
    try:
        res = {}
        for pw_on in (False, True):
            run = fused.FusedYOLOv8(model, ctx, pointwise_kernel=pw_on)
            for _ in range(2):
                run(x)
            torch.cuda.synchronize()
            log.clear()
            for _ in range(reps):
                run(x)
            torch.cuda.synchronize()
            for i, (tag, cin, cout, h, e0, e1) in enumerate(log):
                key = "%d->%d @h=%d #%d" % (cin, cout, h, i % (len(log) // reps))
                res.setdefault(key, {}).setdefault(tag, []).append(e0.elapsed_time(e1))
        rows = {k: {t: round(1e3 * sum(v) / len(v), 1) for t, v in d.items()} for k, d in res.items()}
        tot = {t: round(sum(d.get(t, 0.0) for d in rows.values()), 1) for t in ("cudnn+k5", "k6")}
        print(json.dumps({"pointwise_layers_us_inside_the_forward": rows, "sum_us": tot}))
    finally:
        fused.FusedYOLOv8._pw, fused._Conv.raw, fused.FusedYOLOv8._epi = orig_pw, orig_raw, orig_epi


if __name__ == "__main__":
    main()
