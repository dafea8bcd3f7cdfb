"""YOLOv8 forward with the convolutions in PyTorch (cuDNN) and everything between them in libhvb (K5).

Eager PyTorch spends ~2/3 of a YOLOv8m forward outside the convolutions: a broadcast bias add and a
SiLU pass per layer, `.contiguous()` copies of `chunk` views, `torch.cat`, `nn.Upsample`.  Here each
convolution is called without bias and followed by ONE `hvb_bias_act` pass that adds the bias, applies
SiLU, adds the Bottleneck residual and writes the result both where the next convolution reads it and
into its slice of the C2f / Detect concat buffer; the neck's Upsample + Concat never run as passes of
their own either (the producing epilogue writes the 2x2-replicated pixels straight into the concat
buffer); SPPF's four-way concat is one `hvb_concat_nhwc` pass; layer 0 reads K1's NCHW output directly
(`hvb_stem_conv`); the pointwise (1x1) layers where it is faster (few channels, many pixels) skip cuDNN and run as one
`hvb_pointwise_conv` launch (K6: tcgen05 TF32 GEMM with the same epilogue and destinations).

The layer graph is ultralytics 8.3.148 `yolov8.yaml` (the model the reference loads at
hockey/main.py:77 and runs at :179-184), same as hvb.models.yolov8.YOLOv8 whose (conv+bn folded)
weights this runner borrows.  Returns the raw Detect outputs as SplitHeads (box bins [B,64,H_i,W_i] and class logits
[B,nc,H_i,W_i], channels-last).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from .. import _ffi
from .yolov8 import YOLOv8, ConvBnAct, C2f, SPPF, fuse_conv_bn

_SILU_EXACT, _SILU_FAST, _NONE = 1, 4, 0
CL = torch.channels_last


class _Conv:
    """One folded convolution: channels-last weight, bias kept separate for the K5 epilogue."""

    def __init__(self, conv: torch.nn.Conv2d, device):
        self.w = conv.weight.detach().to(device).contiguous(memory_format=CL)
        b = conv.bias.detach() if conv.bias is not None else torch.zeros(conv.out_channels)
        self.b = b.to(device).float().contiguous()
        self.stride, self.padding, self.cout = conv.stride, conv.padding, conv.out_channels
        self.cin = conv.in_channels
        # pointwise layers can run as one K6 launch (tcgen05 GEMM + epilogue).  Routing follows the per-layer times
        # measured INSIDE the YOLOv8m forward (profiles/r01_k6_pointwise.md): K6 wins for c_out <= 96 and (on par) for
        # 192 -> 192; with more input channels one 128 x 96 tile per CTA re-reads X and W from L2 too
        # often (9 TB/s of L2 traffic at 576 -> 192) and cuDNN + K5 stays ahead
        self.pointwise = (tuple(conv.kernel_size) == (1, 1) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (0, 0)
                          and tuple(conv.dilation) == (1, 1) and conv.groups == 1 and self.cin % 32 == 0
                          and (self.cout % 96 == 0 or self.cout % 64 == 0)
                          and (self.cout <= 96 or (self.cout == 192 and self.cin <= 192)))    # only what was measured
        # K6 feeds fp32 bits to the tensor core, which drops the low 13 mantissa bits; the (static) weights are rounded
        # to TF32 here once, to nearest, so only the activations are truncated
        self.w_tf32 = None
        if self.pointwise:
            bits = self.w.reshape(self.cout, self.cin).contiguous().view(torch.int32)
            self.w_tf32 = ((bits + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()

    def raw(self, x):
        y = torch.conv2d(x, self.w, None, self.stride, self.padding)
        if not y.is_contiguous(memory_format=CL):           # K5 kernels address dense NHWC
            y = y.contiguous(memory_format=CL)
        return y


class FusedYOLOv8:
    def __init__(self, model: YOLOv8, ctx, stem_kernel: bool = True, exact_silu: bool = False, pointwise_kernel=None):
        import copy
        m = copy.deepcopy(model).eval()
        if any(isinstance(x.bn, torch.nn.BatchNorm2d) for x in m.modules() if isinstance(x, ConvBnAct)):
            m = fuse_conv_bn(m)
        self.ctx, self.nc, self.m = ctx, int(model.nc), m
        dev = ctx.device
        self.convs = {}
        for mod in m.modules():
            if isinstance(mod, torch.nn.Conv2d):
                self.convs[mod] = _Conv(mod, dev)
        self.stem_w = m.b0.conv.weight.detach().float().cpu().contiguous().numpy()
        self.stem_b = m.b0.conv.bias.detach().float().cpu().contiguous().numpy()
        self.use_stem = stem_kernel and self.stem_w.shape[0] in (16, 32, 48, 64)
        self._lib, self._h = ctx.lib, ctx.handle
        # K6 computes in TF32 like cuDNN's default convolutions; when the caller has turned TF32 off (fp32 convolutions
        # requested) the pointwise layers stay on cuDNN + K5.  None = follow torch.backends.cudnn.allow_tf32 as of now.
        self.use_pointwise = bool(torch.backends.cudnn.allow_tf32) if pointwise_kernel is None else bool(pointwise_kernel)
        import os
        if pointwise_kernel is None and os.environ.get("HVB_K6") is not None:       # A/B switch for bench.py runs
            self.use_pointwise = self.use_pointwise and os.environ["HVB_K6"] not in ("0", "off", "false")
        self.pw_launches = 0
        # SiLU flavour of the epilogue: the approximate-unit version (<= 1e-6 relative error) keeps the pass HBM-bound
        self._silu = _SILU_EXACT if exact_silu else _SILU_FAST
        # measurement hook (bench.py): when set to a list, every epilogue launch is bracketed by CUDA events on the
        # launching stream and logged as (algorithmic bytes, start event, stop event)
        self.epi_log = None

    # ------------------------------------------------------------------ primitives (ctx lock is held by forward)
    # A "dest" says where the LAST epilogue of a block writes: dict(out1=, off1=, out2=, off2=, up2=).  out1 None = in
    # place (dense, for a following convolution); a concat-buffer slice as out1/out2 makes torch.cat / Upsample free.
    def _epi(self, x, bias, act=None, res=None, out1=None, off1=0, out2=None, off2=0, c2b=0, c2n=None, up2=False):
        npix, c = x.shape[0] * x.shape[2] * x.shape[3], x.shape[1]
        if act is None:
            act = self._silu
        if out1 is None:
            out1 = x
        if c2n is None:
            c2n = c
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        uh, uw = (x.shape[2], x.shape[3]) if (up2 and out2 is not None) else (0, 0)
        log = self.epi_log
        if log is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _ffi.check(self._lib.hvb_bias_act(self._h, p(x), p(bias), p(res), npix, c, act, p(out1), out1.shape[1], off1,
                                          p(out2), out2.shape[1] if out2 is not None else 0, off2, c2b, c2n, uh, uw))
        if log is not None:
            e1.record()
            # algorithmic bytes: read x (+ residual), write every destination element once
            nbytes = 4 * npix * (c * (2 + (res is not None)) + (c2n * (4 if uh else 1) if out2 is not None else 0))
            log.append((nbytes, e0, e1))
        return out1

    def _pw_ok(self, k: _Conv, x, off1=0, out2=None, off2=0, c2b=0, c2n=None, up2=False, res=None, **_):
        return (self.use_pointwise and k.pointwise and not up2 and res is None and x.shape[1] == k.cin
                and x.is_contiguous(memory_format=CL) and off1 % 4 == 0 and off2 % 4 == 0 and c2b % 4 == 0
                and (c2n is None or c2n % 4 == 0) and (out2 is None or out2.shape[1] % 4 == 0))

    def _pw(self, k: _Conv, x, act=None, out1=None, off1=0, out2=None, off2=0, c2b=0, c2n=None, up2=False, res=None):
        """K6: the whole Conv (1x1 convolution + bias + SiLU, same destinations as _epi) in one launch."""
        n, _, h, w = x.shape
        if act is None:
            act = self._silu
        if out1 is None:
            out1 = self._buf(n, k.cout, h, w)
        if c2n is None:
            c2n = k.cout
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        _ffi.check(self._lib.hvb_pointwise_conv(self._h, p(x), k.cin, p(k.w_tf32), p(k.b), n * h * w, k.cin, k.cout, act,
                                                p(out1), out1.shape[1], off1, p(out2),
                                                out2.shape[1] if out2 is not None else 0, off2, c2b, c2n if out2 is not None else 0))
        self.pw_launches += 1
        return out1

    def _buf(self, n, c, h, w):
        return torch.empty((n, c, h, w), dtype=torch.float32, device=self.ctx.device, memory_format=CL)

    def _cba(self, mod: ConvBnAct, x, dest=None):
        k = self.convs[mod.conv]
        if self._pw_ok(k, x, **(dest or {})):
            return self._pw(k, x, **(dest or {}))
        return self._epi(k.raw(x), k.b, **(dest or {}))

    def _c2f(self, mod: C2f, x, dest=None):
        c, nb = mod.c, len(mod.m)
        n, _, h, w = x.shape
        cat = self._buf(n, (2 + nb) * c, h, w)
        k = self.convs[mod.cv1.conv]
        y = self._buf(n, c, h, w)
        # cv1: both halves into the concat buffer, second half also dense for the first Bottleneck
        if self._pw_ok(k, x, out2=y, c2b=c, c2n=c):
            self._pw(k, x, out1=cat, off1=0, out2=y, off2=0, c2b=c, c2n=c)
        else:
            self._epi(k.raw(x), k.b, out1=cat, off1=0, out2=y, off2=0, c2b=c, c2n=c)
        for i, bt in enumerate(mod.m):
            a = self._cba(bt.cv1, y)
            k2 = self.convs[bt.cv2.conv]
            r = k2.raw(a)
            res = y if bt.add else None
            if i == nb - 1:
                self._epi(r, k2.b, res=res, out1=cat, off1=(2 + i) * c)
            else:
                y = self._epi(r, k2.b, res=res, out1=r, out2=cat, off2=(2 + i) * c, c2b=0, c2n=c)
        return self._cba(mod.cv2, cat, dest)

    def _sppf(self, mod: SPPF, x, dest=None):
        y0 = self._cba(mod.cv1, x)
        n, c, h, w = y0.shape
        k = mod.m.kernel_size if isinstance(mod.m.kernel_size, int) else mod.m.kernel_size[0]
        if k == 5 and c % 4 == 0 and 2 * h * w * 16 <= 200 * 1024:
            cat = self._buf(n, 4 * c, h, w)             # three cascaded 5x5 pools + concat in one on-chip pass
            _ffi.check(self._lib.hvb_sppf_pool_concat(self._h, C.c_void_p(y0.data_ptr()), n, h, w, c, C.c_void_p(cat.data_ptr())))
        else:
            y = [y0]
            for _ in range(3):
                y.append(mod.m(y[-1]))
            cat = self._cat(y, [0, 0, 0, 0])
        return self._cba(mod.cv2, cat, dest)

    def _cat(self, srcs, shifts):
        n = srcs[0].shape[0]
        h, w = srcs[0].shape[2] << shifts[0], srcs[0].shape[3] << shifts[0]
        out = self._buf(n, sum(t.shape[1] for t in srcs), h, w)
        k = len(srcs)
        ptrs = (C.c_void_p * 4)(*([t.data_ptr() for t in srcs] + [0] * (4 - k)))
        chans = (C.c_int32 * 4)(*([t.shape[1] for t in srcs] + [0] * (4 - k)))
        shs = (C.c_int32 * 4)(*(list(shifts) + [0] * (4 - k)))
        _ffi.check(self._lib.hvb_concat_nhwc(self._h, ptrs, chans, shs, k, n, h, w, C.c_void_p(out.data_ptr())))
        return out

    def _detect(self, feats):
        """Detect: the 64 box-bin channels and the nc class channels stay in their own dense tensors (SplitHeads): the
        bias add runs in place on each convolution output, and K2a's confidence scan reads contiguous class logits."""
        from ..runtime import SplitHeads
        det, boxes, clss = self.m.detect, [], []
        for i, x in enumerate(feats):
            for branch, dst in ((det.cv2[i], boxes), (det.cv3[i], clss)):
                t = self._cba(branch[1], self._cba(branch[0], x))
                k = self.convs[branch[2]]
                dst.append(self._pw(k, t, act=_NONE) if self._pw_ok(k, t) else self._epi(k.raw(t), k.b, act=_NONE))
        return SplitHeads(boxes, clss)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x: float32 [B,3,H,W] NCHW-contiguous (K1's output), H and W multiples of 32."""
        m, ctx = self.m, self.ctx
        n, _, h, w = x.shape
        if h % 32 or w % 32:
            raise ValueError("input height/width must be multiples of 32, got %dx%d" % (h, w))
        cout = lambda mod: mod.cv2.conv.out_channels
        c3, c4, c5, c12 = cout(m.b4), cout(m.b6), cout(m.b9), cout(m.h12)
        c16, c19 = m.h16.conv.out_channels, m.h19.conv.out_channels
        with ctx.lock:
            ctx._enter()
            # the four neck Concat inputs exist only as these buffers: producers write their slices directly
            cat12 = self._buf(n, c5 + c4, h // 16, w // 16)          # [Upsample(p5), p4]
            cat15 = self._buf(n, c12 + c3, h // 8, w // 8)           # [Upsample(h12), p3]
            cat18 = self._buf(n, c16 + c12, h // 16, w // 16)        # [h16(h15), h12]
            cat21 = self._buf(n, c19 + c5, h // 32, w // 32)         # [h19(h18), p5]
            if self.use_stem and x.is_contiguous():
                y = self._buf(n, self.stem_w.shape[0], h // 2, w // 2)
                _ffi.check(self._lib.hvb_stem_conv(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(self.stem_w.ctypes.data),
                                                   C.c_void_p(self.stem_b.ctypes.data), n, h, w, self.stem_w.shape[0],
                                                   C.c_void_p(y.data_ptr())))
            else:
                y = self._cba(m.b0, x.contiguous(memory_format=CL))
            y = self._c2f(m.b2, self._cba(m.b1, y))
            p3 = self._c2f(m.b4, self._cba(m.b3, y), dict(out2=cat15, off2=c12))
            p4 = self._c2f(m.b6, self._cba(m.b5, p3), dict(out2=cat12, off2=c5))
            self._sppf(m.b9, self._c2f(m.b8, self._cba(m.b7, p4)), dict(out1=cat21, off1=c19, out2=cat12, off2=0, up2=True))
            self._c2f(m.h12, cat12, dict(out1=cat18, off1=c16, out2=cat15, off2=0, up2=True))
            h15 = self._c2f(m.h15, cat15)
            self._cba(m.h16, h15, dict(out1=cat18, off1=0))
            h18 = self._c2f(m.h18, cat18)
            self._cba(m.h19, h18, dict(out1=cat21, off1=0))
            h21 = self._c2f(m.h21, cat21)
            return self._detect([h15, h18, h21])
