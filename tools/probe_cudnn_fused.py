"""Probe: can cuDNN's graph API (nvidia-cudnn-frontend) run conv + bias + SiLU (+ residual) as ONE fused kernel on
fp32 NHWC tensors (TF32 math), with strided (concat-slice) inputs / outputs, on this B200?  Prints support, error
vs the torch conv + libhvb K5 epilogue path, and timings.  Decides whether hvb/models/fused.py can drop the
separate epilogue pass.

    python tools/probe_cudnn_fused.py
"""
import json
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))
import cudnn  # noqa: E402
from hvb.runtime import get_context  # noqa: E402

CL = torch.channels_last
F32 = cudnn.data_type.FLOAT


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def nhwc_strides(c_total, h, w):
    """Strides (in elements, NCHW dim order) of a [N,C,H,W] view whose pixels are c_total apart (a concat slice)."""
    return [h * w * c_total, 1, w * c_total, c_total]


def build(handle, n, cin, cout, h, w, k, stride, x_ld, y_ld, residual, act=True):
    pad = k // 2
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g = cudnn.pygraph(handle=handle, io_data_type=F32, intermediate_data_type=F32, compute_data_type=F32)
    X = g.tensor(name="X", dim=[n, cin, h, w], stride=nhwc_strides(x_ld, h, w), data_type=F32)
    W = g.tensor(name="W", dim=[cout, cin, k, k], stride=[cin * k * k, 1, k * cin, cin], data_type=F32)
    B = g.tensor(name="B", dim=[1, cout, 1, 1], stride=[cout, 1, cout, cout], data_type=F32)
    y = g.conv_fprop(image=X, weight=W, padding=[pad, pad], stride=[stride, stride], dilation=[1, 1], compute_data_type=F32)
    y = g.bias(input=y, bias=B)
    if act:                                # SiLU as sigmoid * x (the swish() binding of frontend 1.18 has its defaults mixed up)
        y = g.mul(a=y, b=g.sigmoid(input=y))
    R = None
    if residual:
        R = g.tensor(name="R", dim=[n, cout, oh, ow], stride=nhwc_strides(cout, oh, ow), data_type=F32)
        y = g.add(a=y, b=R)
    y.set_output(True).set_data_type(F32).set_dim([n, cout, oh, ow]).set_stride(nhwc_strides(y_ld, oh, ow))
    g.validate()
    g.build_operation_graph()
    g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    g.check_support()
    g.build_plans()
    return g, X, W, B, R, y, (oh, ow)


def case(handle, ctx, name, n, cin, cout, h, w, k, stride=1, x_ld=None, y_ld=None, residual=False):
    x_ld, y_ld = x_ld or cin, y_ld or cout
    res = {"case": name}
    try:
        g, X, W, B, R, Y, (oh, ow) = build(handle, n, cin, cout, h, w, k, stride, x_ld, y_ld, residual)
        xbuf = torch.randn(n, x_ld, h, w, device="cuda").contiguous(memory_format=CL)
        x = xbuf[:, x_ld - cin:]                                    # channel slice: pixel pitch x_ld
        wt = (torch.randn(cout, cin, k, k, device="cuda") * 0.1).contiguous(memory_format=CL)
        b = torch.randn(cout, device="cuda")
        ybuf = torch.zeros(n, y_ld, oh, ow, device="cuda").contiguous(memory_format=CL)
        y = ybuf[:, y_ld - cout:]
        r = torch.randn(n, cout, oh, ow, device="cuda").contiguous(memory_format=CL) if residual else None
        ws = torch.empty(max(g.get_workspace_size(), 1), dtype=torch.uint8, device="cuda")
        args = {X: x, W: wt, B: b.view(1, -1, 1, 1), Y: y}
        if residual:
            args[R] = r
        run = lambda: g.execute(args, ws, handle=handle)
        run(); torch.cuda.synchronize()
        # reference: torch conv (TF32) + K5 epilogue
        xd = x.contiguous(memory_format=CL)
        ref = torch.nn.functional.silu(torch.conv2d(xd, wt, b, stride, k // 2))
        if residual:
            ref = ref + r
        res["max_err_vs_torch"] = float((y - ref).abs().max()); res["ref_scale"] = float(ref.abs().max())
        res["untouched_slice_ok"] = bool((ybuf[:, : y_ld - cout] == 0).all())
        res["fused_us"] = round(timed(run), 1)

        def torch_k5():
            raw = torch.conv2d(xd, wt, None, stride, k // 2)
            ctx.bias_act(raw, b, "silu", residual=r)
        res["torch_conv_plus_k5_us"] = round(timed(torch_k5), 1)
        res["torch_conv_only_us"] = round(timed(lambda: torch.conv2d(xd, wt, None, stride, k // 2)), 1)
        res["workspace_MB"] = round(g.get_workspace_size() / 1e6, 2)
        try:
            res["plan"] = g.get_plan_name_at_index(0)
        except Exception:
            pass
    except Exception as e:
        res["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
        res["trace"] = traceback.format_exc().splitlines()[-3:]
    print(json.dumps(res), flush=True)


def main():
    torch.backends.cudnn.benchmark = True
    ctx = get_context(0)
    handle = cudnn.create_handle()
    cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)
    print(json.dumps({"cudnn_frontend": cudnn.__version__, "backend": cudnn.backend_version()}))
    n = 32
    case(handle, ctx, "3x3 48->48 @184x320 dense", n, 48, 48, 184, 320, 3)
    case(handle, ctx, "3x3 48->48 @184x320 +residual", n, 48, 48, 184, 320, 3, residual=True)
    case(handle, ctx, "3x3 48->48 @184x320 out->concat slice (ld 192)", n, 48, 48, 184, 320, 3, y_ld=192)
    case(handle, ctx, "3x3 48->48 @184x320 in<-concat slice (ld 192)", n, 48, 48, 184, 320, 3, x_ld=192)
    case(handle, ctx, "1x1 96->96 @184x320 dense", n, 96, 96, 184, 320, 1)
    case(handle, ctx, "1x1 192->96 @184x320 dense (C2f cv2)", n, 192, 96, 184, 320, 1)
    case(handle, ctx, "3x3 s2 48->96 @368x640", n, 48, 96, 368, 640, 3, stride=2)
    case(handle, ctx, "3x3 96->96 @92x160 +residual", n, 96, 96, 92, 160, 3, residual=True)
    case(handle, ctx, "3x3 288->288 @23x40", n, 288, 288, 23, 40, 3)


if __name__ == "__main__":
    main()
