"""Host-side list logic between detection and captioning, restated from ``ref:util/utils.py`` (it needs the OCR
strings, so it stays on the host in this round; SURVEY.md §8a row P2, §8f-2 lists the device version as "next").

Semantics kept exactly (Python floats on float32-rounded ratios, strict comparisons, list order):
``int_box_area`` (ref:util/utils.py:411-415), ``remove_overlap_new`` (:241-319), element construction and the
stable "content is None last" sort (:444-451), caption fill order (:467-469).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def int_box_area(box: Sequence[float], w: int, h: int) -> int:
    x1, y1, x2, y2 = box
    return (int(x2 * w) - int(x1 * w)) * (int(y2 * h) - int(y1 * h))


def _area(b):
    return (b[2] - b[0]) * (b[3] - b[1])


def _inter(b1, b2):
    return max(0, min(b1[2], b2[2]) - max(b1[0], b2[0])) * max(0, min(b1[3], b2[3]) - max(b1[1], b2[1]))


def _iou_star(b1, b2):
    """max(IoU, inter/area1, inter/area2) with the +1e-6 union guard (ref:util/utils.py:259-267)."""
    inter = _inter(b1, b2)
    union = _area(b1) + _area(b2) - inter + 1e-6
    if _area(b1) > 0 and _area(b2) > 0:
        r1, r2 = inter / _area(b1), inter / _area(b2)
    else:
        r1, r2 = 0, 0
    return max(inter / union, r1, r2)


def _is_inside(b1, b2):
    return _inter(b1, b2) / _area(b1) > 0.80   # ref:util/utils.py:269-273


def remove_overlap_new(boxes: List[dict], iou_threshold: float, ocr_bbox: Optional[List[dict]] = None) -> List[dict]:
    filtered: List = []
    if ocr_bbox:
        filtered.extend(ocr_bbox)
    for i, e1 in enumerate(boxes):
        b1 = e1["bbox"]
        valid = True
        for j, e2 in enumerate(boxes):
            b2 = e2["bbox"]
            if i != j and _iou_star(b1, b2) > iou_threshold and _area(b1) > _area(b2):   # keep the smaller box
                valid = False
                break
        if not valid:
            continue
        if ocr_bbox:
            added = False
            labels = ""
            for e3 in ocr_bbox:
                if added:
                    continue
                b3 = e3["bbox"]
                try:
                    if _is_inside(b3, b1):         # OCR box inside the icon: its text labels the icon
                        labels += e3["content"] + " "
                        filtered.remove(e3)
                    elif _is_inside(b1, b3):       # icon inside an OCR box: drop the icon
                        added = True
                        break
                except Exception:                  # the reference swallows errors here (ref:util/utils.py:300-305)
                    continue
            if not added:
                if labels:
                    filtered.append({"type": "icon", "bbox": e1["bbox"], "interactivity": True, "content": labels,
                                     "source": "box_yolo_content_ocr"})
                else:
                    filtered.append({"type": "icon", "bbox": e1["bbox"], "interactivity": True, "content": None,
                                     "source": "box_yolo_content_yolo"})
        else:
            filtered.append(b1)   # reference quirk kept: bare list when there are no OCR boxes (:317-318)
    return filtered


def remove_overlap_fast(boxes: List[dict], iou_threshold: float, ocr_bbox: List[dict]) -> List[dict]:
    """Vectorised ``remove_overlap_new`` (same float64 operations in the same order, so results are identical to the
    Python loops above; tests/test_host_glue_cpu.py checks equality against them and against the reference).
    Requires dict OCR elements (the normal case, ref:util/utils.py:444)."""
    import numpy as np
    n, m = len(boxes), len(ocr_bbox)
    if n == 0:
        return list(ocr_bbox)
    b = np.asarray([e["bbox"] for e in boxes], np.float64).reshape(n, 4)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def inter(p, q):   # p [P,4], q [Q,4] -> [P,Q]
        iw = np.maximum(0, np.minimum(p[:, None, 2], q[None, :, 2]) - np.maximum(p[:, None, 0], q[None, :, 0]))
        ih = np.maximum(0, np.minimum(p[:, None, 3], q[None, :, 3]) - np.maximum(p[:, None, 1], q[None, :, 1]))
        return iw * ih

    it = inter(b, b)
    union = ((area[:, None] + area[None, :]) - it) + 1e-6
    pos = (area[:, None] > 0) & (area[None, :] > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        r1 = np.where(pos, it / area[:, None], 0.0)
        r2 = np.where(pos, it / area[None, :], 0.0)
        iou = np.maximum(np.maximum(it / union, r1), r2)
    bad = (iou > iou_threshold) & (area[:, None] > area[None, :])
    np.fill_diagonal(bad, False)
    valid = ~bad.any(1)
    filtered: List = list(ocr_bbox)
    if m:
        o = np.asarray([e["bbox"] for e in ocr_bbox], np.float64).reshape(m, 4)
        oarea = (o[:, 2] - o[:, 0]) * (o[:, 3] - o[:, 1])
        io = inter(b, o)                                   # [n, m]
        with np.errstate(divide="ignore", invalid="ignore"):
            ocr_in_icon = io / oarea[None, :] > 0.80       # is_inside(ocr, icon)
            icon_in_ocr = io / area[:, None] > 0.80        # is_inside(icon, ocr)
    removed = [False] * m
    if m:
        # the reference walks the OCR boxes in order and breaks at the first one that contains the icon without being
        # inside it (:296-298); labels are collected from the OCR boxes inside the icon BEFORE that point
        stopm = ~ocr_in_icon & icon_in_ocr
        kstop = np.where(stopm.any(1), stopm.argmax(1), m)                       # [n]
        lab = ocr_in_icon & (np.arange(m)[None, :] < kstop[:, None])             # [n, m]
        dropped_v = (kstop < m).tolist()
        has_lab = lab.any(1).tolist()
    valid_l = valid.tolist()
    for i in range(n):
        if not valid_l[i]:
            continue
        labels = ""
        if m:
            if has_lab[i]:
                for k in np.flatnonzero(lab[i]).tolist():
                    labels += ocr_bbox[k]["content"] + " "
                    removed[k] = True
            if dropped_v[i]:
                continue
        filtered.append({"type": "icon", "bbox": boxes[i]["bbox"], "interactivity": True, "content": labels or None,
                         "source": "box_yolo_content_ocr" if labels else "box_yolo_content_yolo"})
    if m and any(removed):
        # list.remove() drops the FIRST equal element; OCR elements are distinct dicts in practice, handled exactly
        out, pending = [], [ocr_bbox[k] for k in range(m) if removed[k]]
        for e in filtered:
            hit = next((j for j, p in enumerate(pending) if p == e), None) if (pending and e.get("type") == "text") else None
            if hit is not None:
                pending.pop(hit)
                continue
            out.append(e)
        filtered = out
    return filtered


def build_elements(xyxy_ratio: List[List[float]], ocr_ratio: Optional[List[List[float]]], ocr_text: Sequence[str],
                   w: int, h: int, iou_threshold: float, fast: bool = True) -> Tuple[List[dict], int]:
    """ref:util/utils.py:444-451 -> (filtered_boxes_elem sorted with content-None last, starting_idx)."""
    if ocr_ratio is None:
        # ref:util/utils.py:437-444 raises TypeError here (zip over None); the drop-in accepts "no OCR" as an empty list
        ocr_ratio = []
    ocr_elem = [{"type": "text", "bbox": box, "interactivity": False, "content": txt, "source": "box_ocr_content_ocr"}
                for box, txt in zip(ocr_ratio, ocr_text) if int_box_area(box, w, h) > 0]
    icon_elem = [{"type": "icon", "bbox": box, "interactivity": True, "content": None}
                 for box in xyxy_ratio if int_box_area(box, w, h) > 0]
    if fast and ocr_elem:
        filtered = remove_overlap_fast(icon_elem, iou_threshold, ocr_elem)
    else:
        filtered = remove_overlap_new(icon_elem, iou_threshold, ocr_elem)
    if not ocr_elem:
        # normalise the bare-list quirk so downstream code sees dicts
        filtered = [{"type": "icon", "bbox": b, "interactivity": True, "content": None, "source": "box_yolo_content_yolo"}
                    if not isinstance(b, dict) else b for b in filtered]
    elems = sorted(filtered, key=lambda x: x["content"] is None)
    starting_idx = next((i for i, e in enumerate(elems) if e["content"] is None), -1)
    return elems, starting_idx


# ------------------------------------------------------------------------------------------ device overlap filter, host half
# b2p_overlap_filter (csrc/overlap_filter.cu) evaluates the geometry of remove_overlap_new on the GPU and returns flags; the
# strings live here.  ocr_elements() is what the reference builds at ref:util/utils.py:444; elements_from_flags() turns the
# flags back into the reference's sorted element list (:446-451) with the same list semantics as the loops above.
OCR_DEVICE_CAP = 256   # OCR boxes per screenshot the device filter takes (b2p_overlap_filter max_ocr); more -> host path


def ocr_elements(ocr_ratio: Optional[List[List[float]]], ocr_text: Sequence[str], w: int, h: int) -> List[dict]:
    return [{"type": "text", "bbox": box, "interactivity": False, "content": txt, "source": "box_ocr_content_ocr"}
            for box, txt in zip(ocr_ratio or [], ocr_text) if int_box_area(box, w, h) > 0]


def elements_from_flags(icon_ratio: Sequence[Sequence[float]], state: Sequence[int], label_mask, ocr_elem: List[dict],
                        ocr_removed: Sequence[int]) -> List[dict]:
    """icon_ratio [n][4] (fp32 values), state [n] (0 dropped / 1 needs a caption / 2 labelled by OCR), label_mask [n][words]
    (bit k: OCR element k labels the icon), ocr_removed [m]  ->  filtered_boxes_elem, "content is None" last (stable)."""
    m = len(ocr_elem)
    kept_ocr: List[dict] = list(ocr_elem)
    if m and any(ocr_removed[k] for k in range(m)):
        # list.remove() drops the FIRST equal element (ref:util/utils.py:296); equal OCR dicts are handled like the loops do
        pending = [ocr_elem[k] for k in range(m) if ocr_removed[k]]
        out = []
        for e in kept_ocr:
            hit = next((j for j, p_ in enumerate(pending) if p_ == e), None) if pending else None
            if hit is not None:
                pending.pop(hit)
                continue
            out.append(e)
        kept_ocr = out
    labelled, plain = [], []
    for i, st in enumerate(state):
        if st == 2:
            words = label_mask[i]
            labels = "".join(ocr_elem[k]["content"] + " " for k in range(m) if (int(words[k >> 5]) >> (k & 31)) & 1)
            labelled.append({"type": "icon", "bbox": [float(v) for v in icon_ratio[i]], "interactivity": True, "content": labels,
                             "source": "box_yolo_content_ocr"})
        elif st == 1:
            plain.append({"type": "icon", "bbox": [float(v) for v in icon_ratio[i]], "interactivity": True, "content": None,
                          "source": "box_yolo_content_yolo"})
    return kept_ocr + labelled + plain


def fill_captions(elems: List[dict], captions: List[str]) -> None:
    """ref:util/utils.py:467-469: captions are consumed in order by the content-None elements."""
    it = iter(captions)
    for e in elems:
        if e["content"] is None:
            e["content"] = next(it)
