"""K2a parity: head decode vs the restated Detect._inference (CPU torch fp32), NMS keep-set vs the
REAL torchvision.ops.nms (margin-free on identical inputs), and the fused decode+NMS+scale_boxes vs
the restated ultralytics post-process on planted-box head tensors."""
import numpy as np
import pytest
import torch
import torchvision

from hvb import _ffi
from hvb.synth import planted_head, random_boxes
from oracle import ultralytics_restated as ur

pytestmark = pytest.mark.gpu


def make_heads(seed, batch, hw, nc, n_gt=12, dup=3, noise_conf=None):
    rng = np.random.default_rng(seed)
    H, W = hw
    level_hw = [(H // s, W // s) for s in (8, 16, 32)]
    per_img = []
    for _ in range(batch):
        cx, cy = rng.uniform(80, W - 80, n_gt), rng.uniform(80, H - 80, n_gt)
        bw, bh = rng.uniform(20, 110, n_gt), rng.uniform(40, 250, n_gt)
        gt = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        per_img.append(planted_head(rng, level_hw, nc, gt, rng.integers(0, nc, n_gt), dup=dup))
    levels = [torch.from_numpy(np.stack([p[i] for p in per_img])) for i in range(3)]
    return levels


def meta_for(batch, img1, img0):
    gain, px, py = ur.scale_boxes_geometry(img1, img0)
    m = np.zeros((batch,), _ffi.IMG_META)
    m["gain"], m["pad_x"], m["pad_y"] = gain, px, py
    m["clip_w"], m["clip_h"] = img0[1], img0[0]
    m["out_slot"] = np.arange(batch)
    return m


@pytest.mark.parametrize("hw,nc", [((384, 640), 1), ((736, 1280), 2), ((640, 640), 3)])
def test_decode_matches_restated_detect(ctx, hw, nc):
    levels = make_heads(1, 2, hw, nc)
    ref = ur.decode_head(levels, nc).numpy()
    got = ctx.decode_only([l.cuda() for l in levels], nc).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got[:, :4] - ref[:, :4]).max() <= 1e-3            # px
    assert np.abs(got[:, 4:] - ref[:, 4:]).max() <= 1e-6            # confidence


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 257, 1000, 1024, 1025, 3000])
@pytest.mark.parametrize("agnostic", [False, True])
def test_nms_keepset_identical_to_torchvision_margin_free(ctx, n, agnostic):
    rng = np.random.default_rng(n + 7 * agnostic)
    boxes, scores = random_boxes(rng, n, 640, 640, 20, 160)
    if n > 4:                                    # exact duplicates and score ties
        boxes[1] = boxes[0]
        scores[3] = scores[2]
    cls = rng.integers(0, 3, n).astype(np.int32)
    tb, ts = torch.from_numpy(boxes), torch.from_numpy(scores)
    off = torch.zeros(n, 1) if agnostic else torch.from_numpy(cls).float().view(-1, 1) * 7680
    for thr in (0.45, 0.7):
        ref = torchvision.ops.nms(tb + off, ts, thr)[:300].numpy() if n else np.zeros(0, np.int64)
        got = ctx.nms_f32(tb.cuda(), ts.cuda(), torch.from_numpy(cls).cuda(), thr, 300, agnostic).cpu().numpy()
        assert np.array_equal(got, ref), (n, thr, len(got), len(ref))


def test_nms_max_det_truncation(ctx):
    rng = np.random.default_rng(5)
    boxes, scores = random_boxes(rng, 2000, 4000, 4000, 10, 30)
    tb, ts = torch.from_numpy(boxes), torch.from_numpy(scores)
    ref = torchvision.ops.nms(tb, ts, 0.7)
    assert len(ref) > 300
    got = ctx.nms_f32(tb.cuda(), ts.cuda(), None, 0.7, 300, True).cpu().numpy()
    assert np.array_equal(got, ref[:300].numpy())


@pytest.mark.parametrize("hw,img0,nc,conf", [((736, 1280), (1080, 1920), 2, 0.4), ((384, 640), (720, 1280), 1, 0.4),
                                             ((640, 640), (640, 640), 1, 0.25), ((640, 640), (624, 640), 2, 0.4)])
def test_fused_decode_nms_scale_matches_restated_ultralytics(ctx, hw, img0, nc, conf):
    batch = 3
    levels = make_heads(11, batch, hw, nc)
    ref = ur.predict_from_head(levels, nc, hw, [img0] * batch, conf, iou=0.7, max_det=300)
    xyxy, cf, cl, cnt = ctx.decode_nms([l.cuda() for l in levels], nc, conf, 0.7, 300, False, meta=meta_for(batch, hw, img0))
    cnt = cnt.cpu().numpy()
    for b in range(batch):
        rx, rc, rk = ref[b]
        assert cnt[b] == len(rx) > 0
        assert np.array_equal(cl[b, :cnt[b]].cpu().numpy(), rk)                       # same keep-set, same order
        assert np.abs(cf[b, :cnt[b]].cpu().numpy() - rc).max() <= 1e-6
        assert np.abs(xyxy[b, :cnt[b]].cpu().numpy() - rx).max() <= 1e-3


def test_empty_and_random_init_heads(ctx):
    """Random-init YOLO emits nothing above conf=0.4: counts are 0, nothing is written."""
    from hvb.models import build_yolov8
    m = build_yolov8("n", 1)
    with torch.no_grad():
        levels = m(torch.rand(2, 3, 384, 640))
    ref = ur.predict_from_head(levels, 1, (384, 640), [(720, 1280)] * 2, 0.4)
    assert all(len(r[0]) == 0 for r in ref)
    *_, cnt = ctx.decode_nms([l.cuda().contiguous() for l in levels], 1, 0.4, 0.7, 300, False, meta=meta_for(2, (384, 640), (720, 1280)))
    assert (cnt.cpu().numpy() == 0).all()


def test_stress_low_conf_uses_large_tier(ctx):
    """conf=1e-3 on a noisy head: thousands of candidates -> the 8192-candidate tier; still the
    same keep-set as the restated path (margins: planted duplicates dominate)."""
    levels = make_heads(3, 2, (640, 640), 1, n_gt=40, dup=4)
    for l in levels:                       # lift the background so ~25% of anchors pass conf=1e-3
        l[:, 64:] += 0.9 * (torch.rand_like(l[:, 64:]) > 0.75).float()
    conf = 1.2e-3
    dec = ur.decode_head(levels, 1)
    ncand = int((dec[:, 4] > conf).sum(1).max())
    assert 1024 < ncand <= 8192, ncand
    ref = ur.predict_from_head(levels, 1, (640, 640), [(640, 640)] * 2, conf)
    xyxy, cf, cl, cnt = ctx.decode_nms([l.cuda() for l in levels], 1, conf, 0.7, 300, False, meta=meta_for(2, (640, 640), (640, 640)))
    cnt = cnt.cpu().numpy()
    for b in range(2):
        assert cnt[b] == len(ref[b][0])
        assert np.abs(cf[b, :cnt[b]].cpu().numpy() - ref[b][1]).max() <= 1e-6
        assert np.abs(xyxy[b, :cnt[b]].cpu().numpy() - ref[b][0]).max() <= 1e-3


def test_nms_idempotent_property(ctx):
    rng = np.random.default_rng(9)
    boxes, scores = random_boxes(rng, 4000, 1920, 1080, 20, 120)
    tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    k1 = ctx.nms_f32(tb, ts, None, 0.5, 1024, True)
    k2 = ctx.nms_f32(tb[k1.long()], ts[k1.long()], None, 0.5, 1024, True)
    assert np.array_equal(k2.cpu().numpy(), np.arange(len(k1)))


def test_channels_last_heads_give_identical_results(ctx):
    """K2a takes element strides: NHWC (channels_last) head tensors must decode to the same bits."""
    levels = make_heads(21, 2, (736, 1280), 2)
    meta = meta_for(2, (736, 1280), (1080, 1920))
    a = ctx.decode_nms([l.cuda() for l in levels], 2, 0.4, 0.7, 300, False, meta=meta)
    b = ctx.decode_nms([l.cuda().contiguous(memory_format=torch.channels_last) for l in levels], 2, 0.4, 0.7, 300, False, meta=meta)
    cnt = a[3].cpu().numpy()
    assert np.array_equal(cnt, b[3].cpu().numpy()) and cnt.min() > 0
    for i in range(2):
        for x, y in zip(a[:3], b[:3]):
            assert torch.equal(x[i, :cnt[i]], y[i, :cnt[i]])


def test_split_heads_layout_gives_identical_results(ctx):
    """hvb_decode_nms_split (box bins and class logits in separate tensors, class logits dense channels-last — what the
    K5 runner writes) == hvb_decode_nms on the combined [B, 64+nc, H, W] tensors, bit for bit, including the retry tier."""
    from hvb.runtime import SplitHeads
    levels = [l.cuda() for l in make_heads(11, 3, (736, 1280), 2)]
    meta = meta_for(3, (736, 1280), (1080, 1920))
    want = ctx.decode_nms(levels, 2, 0.4, 0.7, 300, False, meta=meta)
    split = SplitHeads([l[:, :64].contiguous(memory_format=torch.channels_last) for l in levels],
                       [l[:, 64:].contiguous(memory_format=torch.channels_last) for l in levels])
    meta_d = ctx.struct_to_device(meta)
    out = (ctx.empty((3, 300, 4), torch.float32), ctx.empty((3, 300), torch.float32), ctx.empty((3, 300), torch.int32),
           torch.zeros((3,), dtype=torch.int32, device="cuda"))
    with ctx.lock:
        ctx._enter()
        ctx.decode_nms_call(split, 2, 0.4, 0.7, 300, False, meta_d, *out)
    k = want[3].cpu().numpy()
    assert np.array_equal(out[3].cpu().numpy(), k) and k.min() > 0
    for a, b in zip(out[:3], want[:3]):
        for i in range(3):
            assert torch.equal(a[i, :k[i]], b[i, :k[i]])
    # the combined view of SplitHeads is the plain Detect output
    for c, l in zip(split, levels):
        assert torch.equal(c, l)
    # retry tier on a list of images
    with ctx.lock:
        ctx._enter()
        out[3].zero_()
        ctx.decode_nms_call(split, 2, 0.4, 0.7, 300, False, meta_d, *out, images_dev=torch.tensor([2, 0], dtype=torch.int32, device="cuda"), n_images=2)
    got = out[3].cpu().numpy()
    assert got[0] == k[0] and got[2] == k[2] and got[1] == 0
