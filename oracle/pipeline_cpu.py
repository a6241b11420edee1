"""ORACLE (test infrastructure): CPU restatement of the whole parse path, ``get_som_labeled_img`` without the overlay
drawing (ref:util/utils.py:417-476), on the seeded stand-ins: PIL LANCZOS letterbox -> fp32 PyTorch YOLOv9-E ->
decode / filter / torchvision NMS -> overlap filter -> cv2 crop+resize loop -> HF Florence-2 greedy generate.

Used by tests (as the checker), by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / ``--impl
reference`` legs (as the thing timed on the host cores).  Never imported by the product package.

Caption mode: the reference's CUDA branch semantics (``do_resize=False``, 64x64 crops, ref:util/utils.py:121) by
default -- the same computation the GPU path performs ("mode-matched"); ``caption_768=True`` restates the CPU
branch (resize to 768x768 bicubic first, :123), a ~48x larger computation (SURVEY.md finding 3).
"""
from __future__ import annotations

import time
from typing import List, Sequence

import numpy as np
import torch

from standin import florence as FS
from . import ref_restate as R


# --------------------------------------------------------------------------- ref:util/utils.py:241-319, 411-415, 444-451
def _int_box_area(box, w, h):
    x1, y1, x2, y2 = box
    ib = [int(x1 * w), int(y1 * h), int(x2 * w), int(y2 * h)]
    return (ib[2] - ib[0]) * (ib[3] - ib[1])


def _remove_overlap(boxes, iou_threshold, ocr_bbox):
    def area(b):
        return (b[2] - b[0]) * (b[3] - b[1])

    def inter(a, b):
        return max(0, min(a[2], b[2]) - max(a[0], b[0])) * max(0, min(a[3], b[3]) - max(a[1], b[1]))

    def iou(a, b):
        it = inter(a, b)
        un = area(a) + area(b) - it + 1e-6
        r1, r2 = (it / area(a), it / area(b)) if area(a) > 0 and area(b) > 0 else (0, 0)
        return max(it / un, r1, r2)

    def inside(a, b):
        return inter(a, b) / area(a) > 0.80

    out = list(ocr_bbox) if ocr_bbox else []
    for i, e1 in enumerate(boxes):
        b1 = e1["bbox"]
        if any(i != j and iou(b1, e2["bbox"]) > iou_threshold and area(b1) > area(e2["bbox"]) for j, e2 in enumerate(boxes)):
            continue
        labels, dropped = "", False
        for e3 in (ocr_bbox or []):
            b3 = e3["bbox"]
            if inside(b3, b1):
                labels += e3["content"] + " "
                if e3 in out:
                    out.remove(e3)
            elif inside(b1, b3):
                dropped = True
                break
        if dropped:
            continue
        out.append({"type": "icon", "bbox": b1, "interactivity": True, "content": labels or None,
                    "source": "box_yolo_content_ocr" if labels else "box_yolo_content_yolo"})
    return out


def build_elements(xyxy_ratio, ocr_ratio, ocr_text, w, h, iou_threshold):
    ocr_elem = [{"type": "text", "bbox": b, "interactivity": False, "content": t, "source": "box_ocr_content_ocr"}
                for b, t in zip(ocr_ratio or [], ocr_text) if _int_box_area(b, w, h) > 0]
    icon_elem = [{"type": "icon", "bbox": b, "interactivity": True, "content": None} for b in xyxy_ratio if _int_box_area(b, w, h) > 0]
    elems = sorted(_remove_overlap(icon_elem, iou_threshold, ocr_elem), key=lambda e: e["content"] is None)
    return elems


# --------------------------------------------------------------------------- the pipeline
class OraclePipeline:
    def __init__(self, yolo=None, florence=None):
        from standin.yolo_weights import yolo_standin
        self.yolo = yolo if yolo is not None else yolo_standin(0)
        self.florence = florence if florence is not None else FS.florence_standin(0)

    @torch.no_grad()
    def detect(self, img_u8: np.ndarray, conf: float, iou: float = 0.1, imgsz=640, max_det=300):
        """ref:util/yolov9.py:115-136 on the fp32 oracle network."""
        H, W = img_u8.shape[:2]
        canvas, scale, pl, pt = R.letterbox_pil(img_u8, imgsz)
        x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
        scores, boxes = R.decode_heads(self.yolo(x))
        b, s, c = R.filter_candidates(scores[0], boxes[0], conf, scale, pl, pt)
        keep, kb, ks = R.nms_and_clamp(b, s, c, iou, max_det, W, H)
        return kb, ks

    @torch.no_grad()
    def caption_ids(self, crops_u8: np.ndarray, max_new_tokens=20, caption_768=False, batch_size=128) -> torch.Tensor:
        """ref:util/utils.py:116-130 (greedy Florence-2 over the crops, in batches of ``batch_size``)."""
        out = []
        for i in range(0, len(crops_u8), batch_size):
            c = torch.from_numpy(np.ascontiguousarray(crops_u8[i:i + batch_size]))
            if caption_768:
                import torch.nn.functional as F
                # CLIP image processor default branch: bicubic resize to 768x768 on the u8 image, then rescale/normalise
                from PIL import Image
                c = torch.from_numpy(np.stack([np.asarray(Image.fromarray(a.numpy()).resize((768, 768), Image.Resampling.BICUBIC)) for a in c]))
                n_img = 577
            else:
                n_img = 5
            pv = FS.pixel_values_from_u8(c)
            ids = FS.input_ids_for(len(c), n_img)
            out.append(self.florence.generate(input_ids=ids, pixel_values=pv, max_new_tokens=max_new_tokens, num_beams=1, do_sample=False))
        width = max(o.shape[1] for o in out)
        out = [torch.nn.functional.pad(o, (0, width - o.shape[1]), value=1) for o in out]
        return torch.cat(out, 0)

    def parse(self, img_u8: np.ndarray, ocr_text: Sequence[str], ocr_bbox, BOX_TRESHOLD=0.01, iou_threshold=0.9,
              imgsz=640, max_new_tokens=20, caption_768=False, timings: dict | None = None, det_boxes=None):
        """-> (filtered_boxes_elem with content = token-id tags, caption ids)."""
        import cv2
        H, W = img_u8.shape[:2]
        t0 = time.perf_counter()
        kb = torch.as_tensor(det_boxes, dtype=torch.float32).reshape(-1, 4) if det_boxes is not None else self.detect(img_u8, BOX_TRESHOLD, 0.1, imgsz)[0]
        t1 = time.perf_counter()
        whwh = torch.Tensor([W, H, W, H])
        xyxy = (kb / whwh).tolist()
        oratio = (torch.tensor(ocr_bbox) / whwh).tolist() if ocr_bbox else None
        elems = build_elements(xyxy, oratio, ocr_text, W, H, iou_threshold)
        boxes = torch.tensor([e["bbox"] for e in elems if e["content"] is None], dtype=torch.float32).reshape(-1, 4)
        crops = []
        for (xa, ya, xb, yb) in R.crop_boxes_int(boxes, W, H):
            crops.append(cv2.resize(img_u8[ya:yb, xa:xb, :], (64, 64)))   # ref:util/utils.py:101-102
        t2 = time.perf_counter()
        ids = self.caption_ids(np.stack(crops), max_new_tokens, caption_768) if crops else torch.zeros((0, 1), dtype=torch.long)
        t3 = time.perf_counter()
        k = 0
        for e in elems:
            if e["content"] is None:
                e["content"] = " ".join(f"<{t}>" for t in ids[k].tolist() if t not in (0, 1, 2))
                k += 1
        if timings is not None:
            timings.update(detect_s=t1 - t0, glue_s=t2 - t1, caption_s=t3 - t2, n_boxes=len(kb), n_crops=len(crops))
        return elems, ids
