"""CPU restatement of the ultralytics 8.3.148 predict path the reference reaches from
``hockey/main.py:179-184`` (``self.player_model(frame, imgsz=1280, conf=0.4, verbose=False)[0]``)
and, for the puck path, from the documented ``sv.InferenceSlicer`` callback (README.md:25,
CLAUDE.md:55).  TEST INFRASTRUCTURE — see oracle/__init__.py.

PARITY UNPINNED: ultralytics is a third-party dependency that is neither vendored under
/root/reference nor installed in this image (version evidence: ultralytics 8.3.148 in
notebooks/train_player_detection.ipynb:508,2009).  The functions restate its published
algorithm:

  letterbox_geometry / letterbox     ultralytics/data/augment.py::LetterBox.__call__
  preprocess                         ultralytics/engine/predictor.py::BasePredictor.preprocess
  make_anchors / decode_head         ultralytics/utils/tal.py::make_anchors, dist2bbox;
                                     ultralytics/nn/modules/head.py::Detect._inference; block.py::DFL
  non_max_suppression                ultralytics/utils/ops.py::non_max_suppression (real torchvision.ops.nms)
  scale_boxes / clip_boxes           ultralytics/utils/ops.py::scale_boxes, clip_boxes
  predict_from_head                  ultralytics/models/yolo/detect/predict.py::postprocess +
                                     sv.Detections.from_ultralytics + the mask at hockey/main.py:189-193
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import cv2
import numpy as np
import torch
import torchvision

STRIDES = (8, 16, 32)
REG_MAX = 16
MAX_WH = 7680
MAX_NMS = 30000


# ------------------------------------------------------------------ LetterBox
def letterbox_geometry(h: int, w: int, imgsz: int | Tuple[int, int] = 640, auto: bool = True,
                       stride: int = 32, scaleup: bool = True):
    """Returns dict(new_unpad=(w,h), top,bottom,left,right, out_h,out_w, r).  Python round() is
    banker's rounding, as in the library."""
    new_h, new_w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
    r = min(new_h / h, new_w / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = new_w - new_unpad[0], new_h - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(new_unpad=new_unpad, top=top, bottom=bottom, left=left, right=right,
                out_h=new_unpad[1] + top + bottom, out_w=new_unpad[0] + left + right, r=r)


def letterbox(img: np.ndarray, imgsz=640, auto: bool = True, stride: int = 32) -> np.ndarray:
    """uint8[H,W,3] -> uint8[H',W',3] with the real cv2.resize / cv2.copyMakeBorder (pad 114)."""
    h, w = img.shape[:2]
    g = letterbox_geometry(h, w, imgsz, auto, stride)
    if (w, h) != g["new_unpad"]:
        img = cv2.resize(img, g["new_unpad"], interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, g["top"], g["bottom"], g["left"], g["right"],
                              cv2.BORDER_CONSTANT, value=(114, 114, 114))


def preprocess(imgs: Sequence[np.ndarray]) -> np.ndarray:
    """list of same-shape letterboxed uint8[H',W',3] BGR -> float32[B,3,H',W'] RGB in [0,1]."""
    im = np.stack(imgs)
    im = im[..., ::-1].transpose((0, 3, 1, 2))
    im = np.ascontiguousarray(im)
    t = torch.from_numpy(im).float()
    t /= 255
    return t.numpy()


# ------------------------------------------------------------------ head decode
def make_anchors(level_hw: Sequence[Tuple[int, int]], strides=STRIDES, offset: float = 0.5):
    """anchor points float32[A,2] (x,y in grid units, +0.5) and strides float32[A]; row-major per
    level, levels concatenated 8 -> 16 -> 32."""
    pts, st = [], []
    for (h, w), s in zip(level_hw, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w,), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(st)


def decode_head(levels: Sequence[torch.Tensor], nc: int) -> torch.Tensor:
    """levels[i]: float32[B, 64+nc, H_i, W_i] raw Detect outputs (cat(cv2_i(x), cv3_i(x))).
    Returns float32[B, 4+nc, A] = cat(xywh * stride, sigmoid(cls)) like Detect._inference."""
    B = levels[0].shape[0]
    no = 4 * REG_MAX + nc
    x_cat = torch.cat([xi.reshape(B, no, -1) for xi in levels], 2)
    box, cls = x_cat.split((4 * REG_MAX, nc), 1)
    anchors, strides = make_anchors([tuple(xi.shape[2:]) for xi in levels])
    A = box.shape[2]
    # DFL: softmax over the 16 bins of each side, expectation with weights 0..15
    d = box.view(B, 4, REG_MAX, A).transpose(2, 1).softmax(1)            # [B,16,4,A]
    wts = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
    dist = (d * wts).sum(1)                                              # [B,4,A]  (l,t,r,b)
    lt, rb = dist.chunk(2, 1)
    anc = anchors.t().unsqueeze(0)                                       # [1,2,A]
    x1y1 = anc - lt
    x2y2 = anc + rb
    c_xy = (x1y1 + x2y2) / 2
    wh = x2y2 - x1y1
    dbox = torch.cat((c_xy, wh), 1) * strides.view(1, 1, A)
    return torch.cat((dbox, cls.sigmoid()), 1)


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        agnostic: bool = False, max_det: int = 300, nc: int = 0) -> List[torch.Tensor]:
    """prediction float32[B, 4+nc, A] -> per image float32[n,6] (x1,y1,x2,y2,conf,cls), score order.
    (classes=None, multi_label=False, no masks; the wall-clock time_limit break is omitted.)"""
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    xc = prediction[:, 4:4 + nc].amax(1) > conf_thres
    prediction = prediction.transpose(-1, -2)
    prediction = torch.cat((xywh2xyxy(prediction[..., :4]), prediction[..., 4:]), dim=-1)
    output = [torch.zeros((0, 6))] * bs
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            continue
        if n > MAX_NMS:
            x = x[x[:, 4].argsort(descending=True)[:MAX_NMS]]
        c = x[:, 5:6] * (0 if agnostic else MAX_WH)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = torchvision.ops.nms(boxes, scores, iou_thres)
        i = i[:max_det]
        output[xi] = x[i]
    return output


def scale_boxes_geometry(img1_shape, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    return gain, pad_x, pad_y


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape) -> torch.Tensor:
    """Rescale xyxy boxes from the letterboxed shape (H',W') to the original (H,W) and clip."""
    gain, pad_x, pad_y = scale_boxes_geometry(img1_shape, img0_shape)
    boxes = boxes.clone()
    boxes[..., 0] -= pad_x
    boxes[..., 1] -= pad_y
    boxes[..., 2] -= pad_x
    boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    boxes[..., 0] = boxes[..., 0].clamp(0, img0_shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, img0_shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, img0_shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, img0_shape[0])
    return boxes


def predict_from_head(levels: Sequence[torch.Tensor], nc: int, img1_shape, img0_shapes, conf: float,
                      iou: float = 0.7, max_det: int = 300, agnostic: bool = False):
    """Full post-process of raw head tensors for a batch: decode -> NMS -> scale_boxes.
    `img0_shapes`: one (H,W) per image (or a single tuple).  Returns a list of
    (xyxy float32[n,4], conf float32[n], cls int64[n]) in score order."""
    pred = decode_head(levels, nc)
    outs = non_max_suppression(pred, conf, iou, agnostic=agnostic, max_det=max_det, nc=nc)
    if isinstance(img0_shapes[0], int):
        img0_shapes = [tuple(img0_shapes)] * len(outs)
    res = []
    for o, s0 in zip(outs, img0_shapes):
        if o.shape[0]:
            o = o.clone()
            o[:, :4] = scale_boxes(img1_shape, o[:, :4], s0)
        res.append((o[:, :4].numpy().astype(np.float32), o[:, 4].numpy().astype(np.float32),
                    o[:, 5].numpy().astype(np.int64)))
    return res


def greedy_nms_f32(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Pure restatement of torchvision.ops.nms in float32: stable descending
    order, suppress iff IoU > thr, IoU = inter / (area_i + area_j - inter).  Used to cross-check
    the real torchvision kernel in the CPU tests."""
    b = boxes.astype(np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    areas = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    sup = np.zeros(len(b), bool)
    keep = []
    for ii, i in enumerate(order):
        if sup[i]:
            continue
        keep.append(i)
        rest = order[ii + 1:]
        xx1 = np.maximum(b[i, 0], b[rest, 0])
        yy1 = np.maximum(b[i, 1], b[rest, 1])
        xx2 = np.minimum(b[i, 2], b[rest, 2])
        yy2 = np.minimum(b[i, 3], b[rest, 3])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        sup[rest[ovr > np.float32(thr)]] = True
    return np.asarray(keep, np.int64)
