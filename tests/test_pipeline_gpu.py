"""GPU: the integrated parse path (omniparser_b200.utils.parse_screenshots / get_som_labeled_img) against
(a) the golden fixtures of the unmodified reference with the detector output injected (pins overlap filter, crop +
resize and caption exactly: identical elements and greedy ids), and (b) the CPU oracle pipeline fed the GPU
detector's own boxes on a batch of screenshots."""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import __graft_entry__ as ge  # noqa: E402
from omniparser_b200 import synth  # noqa: E402
from omniparser_b200.utils import get_som_labeled_img, parse_screenshots  # noqa: E402
from oracle.pipeline_cpu import OraclePipeline  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"
DEV = torch.device("cuda", 0)


def _same_elements(got, ref, ids_got, ids_ref):
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a["type"] == b["type"] and a["source"] == b["source"] and a["bbox"] == b["bbox"]
        assert a["content"].strip() == b["content"].strip()
    assert ids_got.tolist() == (ids_ref.tolist() if torch.is_tensor(ids_ref) else ids_ref)


@pytest.mark.parametrize("name", ["synth_seed0", "synth_seed3_odd", "synth_seed5_3240x2160"])
def test_after_detection_equals_reference_golden(name):
    det, cmp_ = ge.standin_models(DEV)
    g = json.loads((GOLD / f"{name}.json").read_text())
    w, h = g["case"]["size"]
    img = synth.screenshot(g["case"]["seed"], w, h)
    texts, boxes = synth.ocr_boxes(g["case"]["seed"], w, h)
    (elems, ids), = parse_screenshots([img], det, cmp_, [(texts, boxes)], BOX_TRESHOLD=g["box_threshold"],
                                      iou_threshold=g["iou_threshold"], max_new_tokens=g["max_new_tokens"],
                                      _det_override=[g["det_xyxy"]])
    _same_elements(elems, g["parsed_content_list"], ids, g["caption_ids"])


def test_batch_equals_cpu_oracle_on_gpu_boxes():
    det, cmp_ = ge.standin_models(DEV)
    pipe = OraclePipeline(yolo=torch.nn.Identity())
    seeds = [11, 12, 13]
    imgs = [synth.screenshot(s) for s in seeds]
    ocr = [synth.ocr_boxes(s) for s in seeds]
    out = parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
    res = det.predict_batch(imgs, conf=0.05, iou=0.1)
    for i, (elems, ids) in enumerate(out):
        ref_elems, ref_ids = pipe.parse(imgs[i], ocr[i][0], ocr[i][1], BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8,
                                        det_boxes=res[i].boxes.xyxy.cpu())
        _same_elements(elems, ref_elems, ids, ref_ids)
        assert ids.shape[0] > 10


def test_caption_768_mode_equals_cpu_oracle():
    """The reference's CPU-branch caption semantics (768x768 crops, ref:util/utils.py:123) through the fused parse."""
    det, cmp_ = ge.standin_models(DEV)
    pipe = OraclePipeline(yolo=torch.nn.Identity())
    img = synth.screenshot(21)
    texts, boxes = synth.ocr_boxes(21)
    det_boxes = [[100.0, 120.0, 164.0, 190.0], [800.0, 400.0, 840.0, 436.0], [1500.0, 900.0, 1620.0, 960.0]]
    (elems, ids), = parse_screenshots([img], det, cmp_, [(texts, boxes)], BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=6,
                                      _det_override=[det_boxes], caption_size=768)
    ref_elems, ref_ids = pipe.parse(img, texts, boxes, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=6,
                                    det_boxes=torch.tensor(det_boxes), caption_768=True)
    _same_elements(elems, ref_elems, ids, ref_ids)
    assert 1 <= ids.shape[0] <= 3


def test_get_som_labeled_img_api():
    import base64, io
    from PIL import Image
    det, cmp_ = ge.standin_models(DEV)
    img = Image.fromarray(synth.screenshot(0))
    texts, boxes = synth.ocr_boxes(0)
    enc, coords, elems = get_som_labeled_img(img, det, BOX_TRESHOLD=0.05, output_coord_in_ratio=True, ocr_bbox=boxes,
                                             caption_model_processor=cmp_, ocr_text=texts, iou_threshold=0.7, batch_size=128)
    png = Image.open(io.BytesIO(base64.b64decode(enc)))
    assert png.size == img.size
    assert set(coords) == {str(i) for i in range(len(elems))}
    assert all(set(e) == {"type", "bbox", "interactivity", "content", "source"} for e in elems)
    kinds = [e["content"] is None for e in elems]
    assert not any(kinds)
    srcs = [e["source"] for e in elems]
    assert srcs == sorted(srcs, key=lambda s: s == "box_yolo_content_yolo")   # captioned icons last (ref:util/utils.py:449)
    # the drop-in route: reference-style processor + model.generate on host crops gives the same ids as the fused route
    import cv2
    from oracle import ref_restate as R
    H, W = 1080, 1920
    cb = torch.tensor([e["bbox"] for e in elems if e["source"] == "box_yolo_content_yolo"], dtype=torch.float32)
    crops = [Image.fromarray(cv2.resize(np.asarray(img)[ya:yb, xa:xb], (64, 64))) for (xa, ya, xb, yb) in R.crop_boxes_int(cb, W, H)]
    model, proc = cmp_["model"], cmp_["processor"]
    inputs = proc(images=crops, text=["<CAPTION>"] * len(crops), return_tensors="pt", do_resize=False).to(device=model.device, dtype=torch.float16)
    ids = model.generate(input_ids=inputs["input_ids"], pixel_values=inputs["pixel_values"], max_new_tokens=20, num_beams=1, do_sample=False)
    texts2 = [t.strip() for t in proc.batch_decode(ids, skip_special_tokens=True)]
    assert texts2 == [e["content"] for e in elems if e["source"] == "box_yolo_content_yolo"]


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_pipelined_parser_equals_sequential(lanes):
    """The pipelined schedule (detect of batch i+1 | host list logic of batch i | `lanes` caption batches in flight, each
    on its own stream and plan instance) returns exactly what the one-batch-at-a-time path returns, including the first
    batches (whose plans are built on the fly)."""
    from omniparser_b200.utils import PipelinedParser
    det, cmp_ = ge.standin_models(DEV)
    batches = []
    for b in range(6):
        seeds = [40 + 2 * b, 41 + 2 * b]
        batches.append(([synth.screenshot(s) for s in seeds], [synth.ocr_boxes(s) for s in seeds]))
    ref = [parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8) for imgs, ocr in batches]
    pp = PipelinedParser(det, cmp_, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8, caption_lanes=lanes)
    for rep in range(2):
        got = list(pp.run(iter(batches)))
        assert len(got) == len(ref)
        for bi, (gb, rb) in enumerate(zip(got, ref)):
            for si, ((ge_, gi), (re_, ri)) in enumerate(zip(gb, rb)):
                same_boxes = [e["bbox"] for e in ge_] == [e["bbox"] for e in re_]
                assert same_boxes, f"rep {rep} batch {bi} shot {si}: element boxes differ"
                assert gi.shape == ri.shape and torch.equal(gi, ri), (
                    f"rep {rep} batch {bi} (lane {bi % lanes}) shot {si}: caption ids differ in "
                    f"{int((gi != ri).any(1).sum()) if gi.shape == ri.shape else -1} of {ri.shape[0]} rows")
                assert ge_ == re_


@pytest.mark.parametrize("lanes,group", [(1, 2), (2, 2), (2, 3)])
def test_grouped_captioning_equals_sequential(lanes, group):
    """caption_group > 1: the crops of several batches go through Florence-2 in one pass; per-batch results unchanged."""
    from omniparser_b200.utils import PipelinedParser
    det, cmp_ = ge.standin_models(DEV)
    batches = []
    for b in range(7):
        seeds = [60 + 2 * b, 61 + 2 * b]
        batches.append(([synth.screenshot(s) for s in seeds], [synth.ocr_boxes(s) for s in seeds]))
    ref = [parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8) for imgs, ocr in batches]
    pp = PipelinedParser(det, cmp_, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8, caption_lanes=lanes, caption_group=group)
    for rep in range(2):
        got = list(pp.run(iter(batches)))
        assert len(got) == len(ref)
        for bi, (gb, rb) in enumerate(zip(got, ref)):
            for si, ((ge_, gi), (re_, ri)) in enumerate(zip(gb, rb)):
                assert gi.shape == ri.shape and torch.equal(gi, ri), f"rep {rep} batch {bi} shot {si}"
                assert ge_ == re_


def test_edge_cases_no_boxes_and_no_ocr():
    """Empty inputs the reference either mishandles or never produces: no detection above the threshold (only OCR
    elements come back, no caption launch) and no OCR boxes at all (the reference raises TypeError at
    ref:util/utils.py:444; the drop-in treats it as an empty OCR list)."""
    det, cmp_ = ge.standin_models(DEV)
    img = synth.screenshot(2)
    texts, boxes = synth.ocr_boxes(2)
    (elems, ids), = parse_screenshots([img], det, cmp_, [(texts, boxes)], BOX_TRESHOLD=0.9999999, iou_threshold=0.7, max_new_tokens=8)
    assert ids.shape[0] == 0 and all(e["type"] == "text" for e in elems) and len(elems) == len(texts)
    (elems, ids), = parse_screenshots([img], det, cmp_, [([], None)], BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
    assert len(elems) > 10 and all(e["source"] == "box_yolo_content_yolo" for e in elems) and ids.shape[0] == len(elems)
    # a batch whose images yield different crop counts, one of them zero
    out = parse_screenshots([img, np.full_like(img, 200)], det, cmp_, [([], None), ([], None)], BOX_TRESHOLD=0.3, iou_threshold=0.7,
                            max_new_tokens=8)
    assert len(out) == 2 and out[0][1].shape[0] == len(out[0][0]) and out[1][1].shape[0] == len(out[1][0])


def test_device_overlap_filter_path_equals_host_list_logic(monkeypatch):
    """parse_screenshots / PipelinedParser with the overlap filter on the device (default) return exactly what
    they return with the reference's list logic on the host (B2P_HOST_GLUE), on screenshots whose OCR boxes sit inside icons,
    contain icons, and are plain clutter."""
    from omniparser_b200 import utils as U
    det, cmp_ = ge.standin_models(DEV)
    seeds = [60, 61, 62, 63]
    imgs = [synth.screenshot(s) for s in seeds]
    ocr = []
    probe = det.predict_batch(imgs, conf=0.05, iou=0.1)
    for s, r in zip(seeds, probe):
        texts, boxes = synth.ocr_boxes(s)
        for j, b in enumerate(r.boxes.xyxy.cpu().tolist()[:12]):
            x1, y1, x2, y2 = b
            if j % 3 == 0 and x2 - x1 > 12 and y2 - y1 > 12:      # OCR box inside an icon -> label
                boxes.append([int(x1) + 3, int(y1) + 3, int(x2) - 3, int(y2) - 3]); texts.append(f"in{j}")
            elif j % 3 == 1:                                       # OCR box containing an icon -> icon dropped
                boxes.append([max(0, int(x1) - 20), max(0, int(y1) - 20), int(x2) + 20, int(y2) + 20]); texts.append(f"around{j}")
        ocr.append((texts, boxes))
    monkeypatch.setattr(U, "_HOST_GLUE", True)
    ref = U.parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
    monkeypatch.setattr(U, "_HOST_GLUE", False)
    got = U.parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
    assert any(e["source"] == "box_yolo_content_ocr" for el, _ in ref for e in el)
    for (ge_, gi), (re_, ri) in zip(got, ref):
        assert ge_ == re_ and torch.equal(gi, ri)
    # the pipelined parser, batches of 2 (detector numerics depend on the batch size through the GEMM tiling, so the host-logic
    # reference is taken at the same batch size)
    batches = [(imgs[:2], ocr[:2]), (imgs[2:], ocr[2:]), (imgs[:2], ocr[:2])]
    monkeypatch.setattr(U, "_HOST_GLUE", True)
    ref2 = [x for im, oc in batches for x in U.parse_screenshots(im, det, cmp_, oc, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)]
    monkeypatch.setattr(U, "_HOST_GLUE", False)
    pp = U.PipelinedParser(det, cmp_, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8, caption_lanes=2)
    out = list(pp.run(iter(batches)))
    flat = [x for b in out for x in b]
    assert len(flat) == len(ref2)
    for (ge_, gi), (re_, ri) in zip(flat, ref2):
        assert ge_ == re_ and torch.equal(gi, ri)
