"""ORACLE (test infrastructure): generate tests/golden/*.json by running the UNMODIFIED reference
(``/root/reference/util/utils.py::get_som_labeled_img`` + ``util/yolov9.py::YOLOv9Detector``) on the seeded stand-ins.
Runs only where /root/reference exists (this container): ``python -m oracle.make_golden``.

The reference's caption branch depends on ``model.device.type`` (ref:util/utils.py:120-123).  The goldens pin the
CUDA-branch semantics (64x64 crops, ``do_resize=False``), executed on the CPU in fp32: the stand-in caption model
reports a ``device`` whose ``.type`` is 'cuda' so the unmodified reference code takes that branch, and the stand-in
processor's ``.to(device, dtype)`` keeps fp32 (the reference's fp16 cast is a precision choice, not semantics).
"""
from __future__ import annotations

import json
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
from PIL import Image

from omniparser_b200 import synth

from standin import florence as FS
from .shims import import_reference
from standin.yolo_weights import GOLDEN, yolo_standin
from standin.yolov9e import export_torchscript

CASES = [dict(name="synth_seed0", seed=0, size=(1920, 1080)), dict(name="synth_seed3_odd", seed=3, size=(1919, 1079)),
         dict(name="synth_seed5_3240x2160", seed=5, size=(3240, 2160))]   # geometry of ref:imgs/demo_image.jpg (BASELINE configs[0])
BOX_TRESHOLD, IOU = 0.05, 0.7


class _Batch(dict):
    def to(self, device=None, dtype=None):
        return self


class _Processor:
    """What AutoProcessor('microsoft/Florence-2-base') does for ``do_resize=False`` (needs the network, hence restated)."""

    def __call__(self, images, text, return_tensors="pt", do_resize=True):
        assert do_resize is False, "goldens pin the CUDA-branch (64x64) semantics"
        u8 = torch.from_numpy(np.stack([np.asarray(im) for im in images]))
        return _Batch(input_ids=FS.input_ids_for(len(images)), pixel_values=FS.pixel_values_from_u8(u8))

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(f"<{t}>" for t in row if t not in (0, 1, 2)) for row in ids.tolist()]


class _Model:
    def __init__(self, hf):
        self.hf = hf
        self.config = SimpleNamespace(model_type="florence2", name_or_path="seeded/florence2-standin")
        self.device = SimpleNamespace(type="cuda")
        self.ids = []

    def generate(self, **kw):
        out = self.hf.generate(**kw)
        self.ids.append(out)
        return out


def _overlay_sha(png_b64: str, size) -> str:
    """sha256 of the decoded RGB pixels of the reference's annotated PNG (the PNG byte stream itself is encoder-specific)."""
    import base64, hashlib, io
    im = Image.open(io.BytesIO(base64.b64decode(png_b64))).convert("RGB")
    assert im.size == tuple(size)
    return hashlib.sha256(np.asarray(im).tobytes()).hexdigest()


FACADE = dict(name="facade_seed7", seed=7, size=(1920, 1080))


def facade_golden(path, fl):
    """The reference's own facade, ``util/omniparser.py::Omniparser`` (ref:util/omniparser.py:7-32), run UNMODIFIED: the
    easyocr stand-in returns fixed quads (OCR is outside the hot path), ``get_yolo_model`` loads the TorchScript archive,
    and only ``get_caption_model_processor`` -- which needs the network for microsoft/Florence-2-base, ref:util/utils.py:64 --
    is replaced in the facade's namespace by the seeded stand-in pair."""
    import base64, io, sys
    import easyocr
    ro = __import__("util.omniparser", fromlist=["Omniparser"])
    w, h = FACADE["size"]
    img = synth.screenshot(FACADE["seed"], w, h)
    texts, boxes = synth.ocr_boxes(FACADE["seed"], w, h)
    quads = [([[b[0], b[1]], [b[2], b[1]], [b[2], b[3]], [b[0], b[3]]], t, 0.99) for b, t in zip(boxes, texts)]
    import util.utils as ru
    ru.reader.readtext = lambda image_np, **kw: quads                       # the instance created at ref:util/utils.py:22
    cm = _Model(fl)
    ro.get_caption_model_processor = lambda **kw: {"model": cm, "processor": _Processor()}
    op = ro.Omniparser({"som_model_path": str(path), "caption_model_name": "florence2", "caption_model_path": "seeded/florence2-standin",
                        "BOX_TRESHOLD": BOX_TRESHOLD})
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="PNG")
    png, parsed = op.parse(base64.b64encode(buf.getvalue()).decode("ascii"))
    ids = torch.cat(cm.ids, 0) if cm.ids else torch.zeros((0, 1), dtype=torch.long)
    gold = dict(case=FACADE, config=dict(BOX_TRESHOLD=BOX_TRESHOLD), ocr_text=texts, ocr_bbox=boxes, parsed_content_list=parsed,
                caption_ids=ids.tolist(), overlay_sha256=_overlay_sha(png, (w, h)))
    out = GOLDEN / f"{FACADE['name']}.json"
    out.write_text(json.dumps(gold, default=lambda o: float(o) if isinstance(o, (np.floating,)) else o.tolist()))
    print("wrote", out, len(parsed), "elements,", ids.shape[0], "captions")


# Real screenshots shipped with the reference (ref:imgs/*), parsed with the ScreenSpot-Pro eval's call-site parameters
# (ref:eval/ss_pro_gpt4o_omniv2.py:37-51: draw_bbox_config scaled by max(size)/3200, BOX_TRESHOLD 0.05, iou_threshold 0.7,
# output_coord_in_ratio, image passed as a PATH).  BASELINE configs[0] = imgs/demo_image.jpg; configs[4]'s dataset is not
# on disk, so the same call site is replayed on the reference's own images.  The images travel as fixtures
# (tests/golden/imgs/, byte copies: /root/reference does not exist on the GPU box); OCR boxes are seeded (OCR is outside the path).
REAL_CASES = [dict(name="real_demo_image", file="demo_image.jpg", seed=11), dict(name="real_omni3", file="omni3.jpg", seed=12),
              dict(name="real_excel_rgba", file="excel.png", seed=13), dict(name="real_header_bar_thin", file="header_bar_thin.png", seed=14)]


def eval_draw_config(size):
    r = max(size) / 3200
    return {"text_scale": 0.8 * r, "text_thickness": max(int(2 * r), 1), "text_padding": max(int(3 * r), 1), "thickness": max(int(3 * r), 1)}


def real_goldens(ru, det, fl):
    import shutil
    (GOLDEN / "imgs").mkdir(exist_ok=True)
    for case in REAL_CASES:
        src = Path("/root/reference/imgs") / case["file"]
        dst = GOLDEN / "imgs" / case["file"]
        if not dst.exists():
            shutil.copyfile(src, dst)
        image = Image.open(dst)
        w, h = image.size
        texts, boxes = synth.ocr_boxes(case["seed"], w, h)
        raw = det.predict(image.convert("RGB"), conf=BOX_TRESHOLD, iou=0.1)[0].boxes
        cm = _Model(fl)
        cfg = eval_draw_config(image.size)
        png, coords, parsed = ru.get_som_labeled_img(str(dst), det, BOX_TRESHOLD=BOX_TRESHOLD, output_coord_in_ratio=True, ocr_bbox=boxes,
                                                   draw_bbox_config=cfg, caption_model_processor={"model": cm, "processor": _Processor()},
                                                   ocr_text=texts, use_local_semantics=True, iou_threshold=IOU, scale_img=False, batch_size=128)
        ids = torch.cat(cm.ids, 0) if cm.ids else torch.zeros((0, 1), dtype=torch.long)
        gold = dict(case=dict(case, size=[w, h], mode=image.mode), box_threshold=BOX_TRESHOLD, iou_threshold=IOU, max_new_tokens=20,
                    draw_bbox_config=cfg, ocr_text=texts, ocr_bbox=boxes,
                    det_xyxy=[[float(np.float32(v)) for v in b] for b in raw.xyxy.tolist()], det_conf=[float(c) for c in raw.conf.tolist()],
                    parsed_content_list=parsed, caption_ids=ids.tolist(), label_coordinates=coords,
                    overlay_sha256=_overlay_sha(png, (w, h)))
        out = GOLDEN / f"{case['name']}.json"
        out.write_text(json.dumps(gold, default=lambda o: float(o) if isinstance(o, (np.floating,)) else o.tolist()))
        print("wrote", out, len(raw.xyxy), "boxes,", ids.shape[0], "captions", flush=True)


def main():
    import sys
    ru, ry = import_reference()
    m = yolo_standin(0)
    path = Path("/tmp/b2p_golden/icon_detect_v3/model.pt")
    export_torchscript(m, path, (640, 640))
    det = ru.get_yolo_model(str(path), device="cpu")
    assert type(det).__name__ == "YOLOv9Detector"
    fl = FS.florence_standin(0)
    if "real" in sys.argv[1:]:          # python -m oracle.make_golden real  -> only the real-image goldens
        real_goldens(ru, det, fl)
        return
    for case in CASES:
        w, h = case["size"]
        img = synth.screenshot(case["seed"], w, h)
        texts, boxes = synth.ocr_boxes(case["seed"], w, h)
        raw = det.predict(Image.fromarray(img), conf=BOX_TRESHOLD, iou=0.1)[0].boxes
        cm = _Model(fl)
        png, coords, parsed = ru.get_som_labeled_img(Image.fromarray(img), det, BOX_TRESHOLD=BOX_TRESHOLD, output_coord_in_ratio=True,
                                                   ocr_bbox=boxes, draw_bbox_config=None,
                                                   caption_model_processor={"model": cm, "processor": _Processor()},
                                                   ocr_text=texts, use_local_semantics=True, iou_threshold=IOU, scale_img=False,
                                                   batch_size=128)
        ids = torch.cat(cm.ids, 0) if cm.ids else torch.zeros((0, 1), dtype=torch.long)
        gold = dict(case=case, box_threshold=BOX_TRESHOLD, iou_threshold=IOU, max_new_tokens=20,
                    det_xyxy=[[float(np.float32(v)) for v in b] for b in raw.xyxy.tolist()], det_conf=[float(c) for c in raw.conf.tolist()],
                    parsed_content_list=parsed, caption_ids=ids.tolist(), label_coordinates=coords,
                    overlay_sha256=_overlay_sha(png, (w, h)))
        out = GOLDEN / f"{case['name']}.json"
        out.write_text(json.dumps(gold, default=lambda o: float(o) if isinstance(o, (np.floating,)) else o.tolist()))
        print("wrote", out, len(raw.xyxy), "boxes,", ids.shape[0], "captions")
    facade_golden(path, fl)
    real_goldens(ru, det, fl)


if __name__ == "__main__":
    main()
