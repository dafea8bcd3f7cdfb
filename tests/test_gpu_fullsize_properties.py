"""Size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes): chunk of 32
1080p frames through K1a, 32-image K2a on 736x1280 heads, 768-crop K3b, N=2000 K4a, full-layer K5 — each checked
through an invariant of the operation instead of an element-by-element oracle comparison (those run at small sizes
in the per-kernel test files)."""
import numpy as np
import pytest
import torch

from hvb import _ffi

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def test_k1a_chunk_of_32_frames_is_frame_independent(ctx):
    """Letterboxing a chunk equals letterboxing its frames one by one (bit-exact), and the padding rows are 114/255."""
    rng = np.random.default_rng(0)
    frames = torch.from_numpy(rng.integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).cuda()
    big = ctx.letterbox_plan(32, 1080, 1920, _ffi.LB_WHOLE, 1280)
    one = ctx.letterbox_plan(1, 1080, 1920, _ffi.LB_WHOLE, 1280)
    out = big.class_views(big.run(frames))[0]
    assert out.shape == (32, 3, 736, 1280)
    for i in (0, 13, 31):
        assert torch.equal(out[i], one.class_views(one.run(frames[i:i + 1].contiguous()))[0][0])
    pad = torch.tensor(114.0 / 255.0, dtype=torch.float32, device="cuda")
    assert (out[:, :, :8, :] == pad).all() and (out[:, :, 728:, :] == pad).all()        # 8 rows of padding top and bottom
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0


def test_k2a_32_images_round_trip_of_planted_boxes(ctx):
    """32 x (736x1280, nc=2): every planted box comes back (within 1.5 px after scale_boxes), nothing else does, and
    running NMS again on the survivors keeps all of them."""
    from hvb.synth import planted_head
    from oracle import ultralytics_restated as ur
    rng = np.random.default_rng(1)
    H, W, B, K = 736, 1280, 32, 12
    lv = [(H // s, W // s) for s in (8, 16, 32)]
    gain, px, py = ur.scale_boxes_geometry((H, W), (1080, 1920))
    per, gts = [], []
    for _ in range(B):
        cx, cy = rng.uniform(120, W - 120, K), rng.uniform(120, H - 120, K)
        bw, bh = rng.uniform(30, 70, K), rng.uniform(60, 160, K)
        gt = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        gts.append((gt - np.array([px, py, px, py])) / gain)
        per.append(planted_head(rng, lv, 2, gt, rng.integers(0, 2, K), dup=1, conf_lo=0.5, conf_hi=0.95))
    levels = [torch.from_numpy(np.stack([p[i] for p in per])).cuda() for i in range(3)]
    meta = np.zeros((B,), _ffi.IMG_META)
    meta["gain"], meta["pad_x"], meta["pad_y"], meta["clip_w"], meta["clip_h"], meta["out_slot"] = gain, px, py, 1920, 1080, np.arange(B)
    xyxy, conf, cls, cnt = ctx.decode_nms(levels, 2, 0.4, 0.7, 300, False, meta=meta)
    cnt_h, xyxy_h = cnt.cpu().numpy(), xyxy.cpu().numpy()
    for b in range(B):
        got = xyxy_h[b, :cnt_h[b]]
        d = np.abs(got[:, None, :] - gts[b][None, :, :]).max(-1)          # [kept, K]
        assert (d.min(0) < 1.5).all(), "a planted box was lost"
        assert (d.min(1) < 1.5).all(), "a detection that was never planted"
        assert cnt_h[b] <= K + 2
    b0 = 5
    keep = ctx.nms_f32(xyxy[b0, :cnt_h[b0]].contiguous(), conf[b0, :cnt_h[b0]].contiguous(), cls[b0, :cnt_h[b0]].contiguous(), 0.7)
    assert len(keep) == cnt_h[b0]                                         # idempotent


def test_k3b_flat_colour_crops_give_constant_planes(ctx):
    """768 crops cut from flat-colour 1080p frames: any interpolation of a constant is that constant, so every output
    plane equals ((v/255) - mean_c) / std_c exactly, whatever the ROI geometry."""
    rng = np.random.default_rng(2)
    nf, per = 64, 12
    cols = rng.integers(0, 256, (nf, 3), dtype=np.uint8)
    frames = torch.from_numpy(np.broadcast_to(cols[:, None, None, :], (nf, 1080, 1920, 3)).copy()).cuda()
    x0 = rng.uniform(0, 1700, nf * per); y0 = rng.uniform(0, 800, nf * per)
    boxes = np.stack([x0, y0, x0 + rng.uniform(40, 110, nf * per), y0 + rng.uniform(100, 250, nf * per)], 1).astype(np.float32)
    fidx = np.repeat(np.arange(nf), per).astype(np.int32)
    cd = ctx.crops_from_boxes(torch.from_numpy(boxes).cuda(), torch.from_numpy(fidx).cuda(), 1080, 1920)
    out, valid = ctx.mnv3_preprocess(frames, cd, nf * per)
    assert bool((valid == 1).all())
    mean = torch.tensor([0.485, 0.456, 0.406]); std = torch.tensor([0.229, 0.224, 0.225])
    want = ((torch.from_numpy(cols[fidx].astype(np.float32)) / 255.0) - mean) / std                # [n, 3]
    o = out.cpu()
    assert torch.equal(o, want[:, :, None, None].expand_as(o).contiguous())


def test_k4a_affinity_invariants_at_n2000(ctx):
    """N=2000, D=625: symmetric, unit diagonal, values in [0,1], tensor-core mode == fp64 mode within 1e-3 relative."""
    rng = np.random.default_rng(3)
    base = rng.normal(0, 1, (500, 625))
    x = np.vstack([base + rng.normal(0, 0.02, base.shape) for _ in range(4)])     # near-duplicates: non-trivial off-diagonals
    xd = torch.from_numpy(x).cuda()
    _, a0 = ctx.gram_affinity(xd, 1.0, 0)
    _, a1 = ctx.gram_affinity(xd, 1.0, 1)
    assert torch.equal(torch.diagonal(a0), torch.ones(2000, dtype=torch.float64, device="cuda"))
    assert float(a0.min()) >= 0.0 and float(a0.max()) <= 1.0
    big = a1 > 1e-300
    assert float(((a0 - a0.T).abs()[big] / a1[big]).max()) <= 1e-3
    assert int(big.sum()) > 2000                                                   # there ARE non-trivial off-diagonal entries
    assert float(((a0[big] - a1[big]).abs() / a1[big]).max()) <= 1e-3


def test_k5_epilogue_on_a_full_layer0_tensor(ctx):
    """32 x 48 x 368 x 640 (1.45 GB, the largest activation of the YOLOv8m forward): act=none is a plain fp32 add
    (bit-exact vs torch on the device); SiLU through the dual-destination path lands identically in both outputs."""
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((32, 48, 368, 640), device="cuda", generator=g).contiguous(memory_format=CL)
    b = torch.randn(48, device="cuda", generator=g)
    ref = x + b.view(1, -1, 1, 1)
    out = ctx.bias_act(x.clone(memory_format=torch.preserve_format), b, "none")
    assert torch.equal(out, ref)
    del ref, out
    cat = torch.zeros((32, 96, 368, 640), device="cuda").contiguous(memory_format=CL)
    y = ctx.bias_act(x, b, "silu", out1=x, out2=cat, out2_off=48, c2_begin=0, c2_count=48)
    assert torch.equal(cat[:, 48:], y) and float(cat[:, :48].abs().max()) == 0.0
    assert float(y.min()) >= -0.2785                                               # min of SiLU
