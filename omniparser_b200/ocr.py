"""OCR pre-step adapter: ``check_ocr_box`` with the reference's signature, return value and box
conversions (ref:util/utils.py:498-549).  The text detectors/recognisers themselves (EasyOCR CRAFT + CRNN, PaddleOCR DB +
SVTR) are a separate model family outside the accelerated path: this module only binds whichever engine is installed -- or
injected with :func:`set_engines` -- lazily (the reference constructs both engines at import time, ref:util/utils.py:20-31,
which makes ``import util.utils`` fail on a machine without them), and raises a clear error when none is available.
"""
from __future__ import annotations

from typing import Union

import numpy as np
from PIL import Image

_ENGINES = {"easyocr": None, "paddle": None}


def set_engines(easyocr_reader=None, paddle_ocr=None) -> None:
    """Inject engine objects: ``easyocr_reader.readtext(image_np, **easyocr_args)`` -> [(quad, text, conf)],
    ``paddle_ocr.ocr(image_np, cls=False)`` -> [[(quad, (text, conf)), ...]]."""
    if easyocr_reader is not None:
        _ENGINES["easyocr"] = easyocr_reader
    if paddle_ocr is not None:
        _ENGINES["paddle"] = paddle_ocr


def _easyocr():
    if _ENGINES["easyocr"] is None:
        try:
            import easyocr
        except ImportError as exc:
            raise RuntimeError("check_ocr_box: EasyOCR is not installed; install it, pass use_paddleocr=True, or inject a reader "
                               "with omniparser_b200.ocr.set_engines(easyocr_reader=...)") from exc
        _ENGINES["easyocr"] = easyocr.Reader(["en"])                  # ref:util/utils.py:21-22
    return _ENGINES["easyocr"]


def _paddle():
    if _ENGINES["paddle"] is None:
        try:
            from paddleocr import PaddleOCR
        except ImportError as exc:
            raise RuntimeError("check_ocr_box: PaddleOCR is not installed; install it or inject an engine with "
                               "omniparser_b200.ocr.set_engines(paddle_ocr=...)") from exc
        _ENGINES["paddle"] = PaddleOCR(lang="en", use_angle_cls=False, use_gpu=False, show_log=False, max_batch_size=1024,
                                       use_dilation=True, det_db_score_mode="slow", rec_batch_num=1024)   # ref:util/utils.py:23-31
    return _ENGINES["paddle"]


def get_xywh(quad):
    x, y, w, h = quad[0][0], quad[0][1], quad[2][0] - quad[0][0], quad[2][1] - quad[0][1]
    return int(x), int(y), int(w), int(h)


def get_xyxy(quad):
    x, y, xp, yp = quad[0][0], quad[0][1], quad[2][0], quad[2][1]
    return int(x), int(y), int(xp), int(yp)


def get_xywh_yolo(box):
    x, y, w, h = box[0], box[1], box[2] - box[0], box[3] - box[1]
    return int(x), int(y), int(w), int(h)


def check_ocr_box(image_source: Union[str, Image.Image], display_img=True, output_bb_format="xywh", goal_filtering=None,
                  easyocr_args=None, use_paddleocr=False):
    """ref:util/utils.py:514-549 -> ``((texts, boxes), goal_filtering)``."""
    if isinstance(image_source, str):
        image_source = Image.open(image_source)
    if image_source.mode == "RGBA":
        image_source = image_source.convert("RGB")
    image_np = np.array(image_source)
    if use_paddleocr:
        text_threshold = 0.5 if easyocr_args is None else easyocr_args["text_threshold"]
        result = _paddle().ocr(image_np, cls=False)[0]
        coord = [item[0] for item in result if item[1][1] > text_threshold]
        text = [item[1][0] for item in result if item[1][1] > text_threshold]
    else:
        result = _easyocr().readtext(image_np, **(easyocr_args or {}))
        coord = [item[0] for item in result]
        text = [item[1] for item in result]
    if display_img:
        import cv2
        opencv_img = cv2.cvtColor(image_np, cv2.COLOR_RGB2BGR)
        bb = []
        for item in coord:
            x, y, a, b = get_xywh(item)
            bb.append((x, y, a, b))
            cv2.rectangle(opencv_img, (x, y), (x + a, y + b), (0, 255, 0), 2)
        try:
            from matplotlib import pyplot as plt
            plt.imshow(cv2.cvtColor(opencv_img, cv2.COLOR_BGR2RGB))
        except ImportError:
            pass
    elif output_bb_format == "xywh":
        bb = [get_xywh(item) for item in coord]
    elif output_bb_format == "xyxy":
        bb = [get_xyxy(item) for item in coord]
    else:
        raise UnboundLocalError(f"output_bb_format {output_bb_format!r}: the reference only defines 'xywh' and 'xyxy'")
    return (text, bb), goal_filtering
