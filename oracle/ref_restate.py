"""ORACLE (test infrastructure): CPU restatement of the reference's detector pre/post-processing.

``/root/reference`` does not exist on the GPU box, so the wrapper logic of
``ref:util/yolov9.py`` is restated here with the same library calls in the same
order (PIL LANCZOS resize + paste, torch fp32 decode, torchvision
``batched_nms``) and, for the integer stages, additionally as pure numpy
(no PIL / cv2) so the arithmetic spec the CUDA kernels implement is pinned
independently.  ``tests/test_oracle_cpu.py`` checks every function here against
the unmodified reference (when ``/root/reference`` is present) and against
Pillow / OpenCV / torchvision themselves.  parity: pinned on Pillow 12.2,
OpenCV 4.13, torchvision 0.26 (the versions in this image); the network
arithmetic itself is "parity unpinned" (see standin/yolov9e.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

STRIDES = (8, 16, 32)


# ----------------------------------------------------------------------------- ref:util/yolov9.py:52-87
def letterbox_geometry(w: int, h: int, imgsz=640):
    """(target_w, target_h, scale, resized_w, resized_h, pad_left, pad_top); ref:util/yolov9.py:52-61,73-80."""
    if isinstance(imgsz, int):
        tw = th = imgsz
    else:
        th, tw = imgsz
    tw = ((int(tw) + 31) // 32) * 32
    th = ((int(th) + 31) // 32) * 32
    scale = min(tw / w, th / h)
    rw, rh = int(w * scale), int(h * scale)
    return tw, th, scale, rw, rh, (tw - rw) // 2, (th - rh) // 2


def letterbox_pil(img_u8: np.ndarray, imgsz=640):
    """u8 HWC -> (u8 canvas [th,tw,3], scale, pad_left, pad_top) through Pillow, as ref:util/yolov9.py:82-84."""
    from PIL import Image

    h, w = img_u8.shape[:2]
    tw, th, scale, rw, rh, pl, pt = letterbox_geometry(w, h, imgsz)
    resized = Image.fromarray(img_u8).resize((rw, rh), Image.Resampling.LANCZOS)
    canvas = Image.new("RGB", (tw, th), (114, 114, 114))
    canvas.paste(resized, (pl, pt))
    return np.asarray(canvas), scale, pl, pt


def _lanczos(x: float) -> float:
    def sinc(v):
        if v == 0.0:
            return 1.0
        v = v * math.pi
        return math.sin(v) / v

    return sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def lanczos_coeffs(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc (22-bit fixed point)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)   # Pillow accumulates left to right in double; Python float sum does the same
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_lanczos_numpy(img_u8: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Two-pass integer FIR (horizontal first, u8 intermediate) = Pillow ImagingResample for 8-bit RGB."""
    h, w, _ = img_u8.shape
    cur = img_u8
    if out_w != w:
        bounds, kk = lanczos_coeffs(w, out_w)
        out = np.empty((h, out_w, 3), np.uint8)
        src = cur.astype(np.int64)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (1 << 21) + (src[:, x0:x0 + n, :] * kk[xx, :n, None].astype(np.int64)[None]).sum(1)
            out[:, xx, :] = np.clip(acc >> 22, 0, 255)
        cur = out
    if out_h != h:
        bounds, kk = lanczos_coeffs(h, out_h)
        out = np.empty((out_h, cur.shape[1], 3), np.uint8)
        src = cur.astype(np.int64)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (1 << 21) + (src[y0:y0 + n] * kk[yy, :n, None, None].astype(np.int64)).sum(0)
            out[yy] = np.clip(acc >> 22, 0, 255)
        cur = out
    return cur


def letterbox_numpy(img_u8: np.ndarray, imgsz=640):
    h, w = img_u8.shape[:2]
    tw, th, scale, rw, rh, pl, pt = letterbox_geometry(w, h, imgsz)
    canvas = np.full((th, tw, 3), 114, np.uint8)
    canvas[pt:pt + rh, pl:pl + rw] = resize_lanczos_numpy(img_u8, rw, rh)
    return canvas, scale, pl, pt


# ----------------------------------------------------------------------------- ref:util/yolov9.py:89-136
def decode_heads(outputs):
    """6 head tensors -> (scores [B,A,nc], boxes [B,A,4] letterboxed xyxy); ref:util/yolov9.py:89-108."""
    logits, boxes = [], []
    for i, stride in enumerate(STRIDES):
        cl, dist = outputs[2 * i], outputs[2 * i + 1]
        b, nc, h, w = cl.shape
        cl = cl.permute(0, 2, 3, 1).reshape(b, -1, nc)
        dist = dist.permute(0, 2, 3, 1).reshape(b, -1, 4) * stride
        gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        anchors = (torch.stack((gx, gy), dim=-1).reshape(-1, 2) + 0.5) * stride
        lt, rb = dist.chunk(2, dim=-1)
        boxes.append(torch.cat((anchors - lt, anchors + rb), dim=-1))
        logits.append(cl)
    return torch.cat(logits, dim=1).sigmoid(), torch.cat(boxes, dim=1)


def filter_candidates(scores, boxes, conf, scale, pad_left, pad_top):
    """class max, strict conf filter, un-letterbox; ref:util/yolov9.py:123-129 (image 0 of the batch)."""
    s, c = scores.max(dim=-1)
    valid = s > conf
    s, c, b = s[valid], c[valid], boxes[valid].clone()
    b[:, [0, 2]] = (b[:, [0, 2]] - pad_left) / scale
    b[:, [1, 3]] = (b[:, [1, 3]] - pad_top) / scale
    return b, s, c


def nms_and_clamp(boxes, scores, cls, iou, max_det, img_w, img_h):
    """torchvision batched_nms, truncate, clamp; ref:util/yolov9.py:131-135. Returns (keep, boxes, scores)."""
    from torchvision.ops import batched_nms

    keep = batched_nms(boxes, scores, cls, iou)[:max_det]
    b, s = boxes[keep].clone(), scores[keep]
    b[:, [0, 2]] = b[:, [0, 2]].clamp(0, img_w)
    b[:, [1, 3]] = b[:, [1, 3]].clamp(0, img_h)
    return keep, b, s


def greedy_nms_numpy(boxes: np.ndarray, scores: np.ndarray, cls: np.ndarray, iou: float, max_det: int) -> np.ndarray:
    """Pure-numpy statement of the NMS rule the kernel implements (fp32 arithmetic, op order of torchvision's
    nms_kernel): sort by score descending, ties by ascending index; suppress j when IoU(i, j) > iou (strict);
    classes separated with the coordinate trick when n <= 1000, by class id otherwise (tv:ops/boxes.py:51-121)."""
    n = len(boxes)
    if n == 0:
        return np.zeros((0,), np.int64)
    boxes = boxes.astype(np.float32)
    trick = n * 4 <= 4000
    if trick:
        off = cls.astype(np.float32) * (boxes.max() + np.float32(1))
        boxes = boxes + off[:, None]
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    x1, y1, x2, y2 = boxes.T
    area = (x2 - x1) * (y2 - y1)
    keep, dead = [], np.zeros(n, bool)
    for _i, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        if len(keep) >= max_det:
            break
        rest = order[_i + 1:]
        w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            ovr = inter / (area[i] + area[rest] - inter)
        sup = ovr.astype(np.float64) > iou
        if not trick:
            sup &= cls[rest] == cls[i]
        dead[rest[sup]] = True
    return np.asarray(keep, np.int64)


# ----------------------------------------------------------------------------- ref:util/utils.py:97-103
def crop_boxes_int(boxes_ratio: torch.Tensor, W: int, H: int):
    """int(coord * shape) on float32 tensors (ref:util/utils.py:99-100)."""
    out = []
    for c in boxes_ratio:
        out.append((int(c[0] * W), int(c[1] * H), int(c[2] * W), int(c[3] * H)))
    return out


def resize_bilinear_cv2_numpy(crop: np.ndarray, out: int = 64) -> np.ndarray:
    """Pure-numpy statement of cv2.resize(crop, (out, out)) for 8UC3 INTER_LINEAR (OpenCV 4.x resize.cpp)."""
    sh, sw, _ = crop.shape
    if sw == 2 * out and sh == 2 * out:   # OpenCV switches to INTER_AREA for exact 2x decimation
        c = crop.astype(np.int32)
        return ((c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def coeffs(src, snap):
        scale = 1.0 / (out / src)
        d = np.arange(out, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = f - s.astype(np.float32)
        if snap:
            lo = s < 0
            f[lo], s[lo] = 0, 0
            hi = s >= src - 1
            f[hi], s[hi] = 0, src - 1
        w1 = np.rint(f * np.float32(2048)).astype(np.int32)
        w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
        return s, w0, w1

    sx, a0, a1 = coeffs(sw, True)
    sy, b0, b1 = coeffs(sh, False)
    src = crop.astype(np.int32)
    x1 = np.minimum(sx + 1, sw - 1)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]   # [sh, out, 3]
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    v = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------- overlap filter flags
def overlap_flags_loops(icon_ratio, ocr_boxes, w, h, iou_threshold):
    """The reference loops of ``remove_overlap_new`` (ref:util/utils.py:241-319) + the ``int_box_area`` filter (:411-415,
    :445), emitting FLAGS instead of element lists -- the contract of the device kernel ``b2p_overlap_filter``
    (omniparser_b200/csrc/overlap_filter.cu): state per icon (0 dropped / 1 needs a caption / 2 labelled by OCR text),
    label bit masks, removed flags of the OCR boxes.  icon_ratio / ocr_boxes: lists of Python-float xyxy ratio boxes
    (OCR boxes already int_box_area-filtered)."""
    def area(b):
        return (b[2] - b[0]) * (b[3] - b[1])

    def inter(b1, b2):
        return max(0, min(b1[2], b2[2]) - max(b1[0], b2[0])) * max(0, min(b1[3], b2[3]) - max(b1[1], b2[1]))

    def iou(b1, b2):
        it = inter(b1, b2)
        union = area(b1) + area(b2) - it + 1e-6
        r1, r2 = (it / area(b1), it / area(b2)) if area(b1) > 0 and area(b2) > 0 else (0, 0)
        return max(it / union, r1, r2)

    n, m = len(icon_ratio), len(ocr_boxes)
    words = max(1, (m + 31) // 32)
    state = [0] * n
    mask = [[0] * words for _ in range(n)]
    removed = [0] * m
    live = [i for i, b in enumerate(icon_ratio)
            if (int(b[2] * w) - int(b[0] * w)) * (int(b[3] * h) - int(b[1] * h)) > 0]
    for i in live:
        b1 = icon_ratio[i]
        if any(j != i and iou(b1, icon_ratio[j]) > iou_threshold and area(b1) > area(icon_ratio[j]) for j in live):
            continue
        dropped, labelled = False, False
        for k, b3 in enumerate(ocr_boxes):
            if inter(b3, b1) / area(b3) > 0.80:
                mask[i][k >> 5] |= 1 << (k & 31)
                removed[k] = 1
                labelled = True
            elif inter(b1, b3) / area(b1) > 0.80:
                dropped = True
                break
        state[i] = 0 if dropped else (2 if labelled else 1)
    return state, mask, removed
