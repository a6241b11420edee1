"""Drop-in for the reference facade ``util/omniparser.py::Omniparser`` (ref:util/omniparser.py:7-32): same config keys,
same ``parse(image_base64) -> (annotated PNG base64, parsed_content_list)``, same draw configuration, with the three hot
functions taken from :mod:`omniparser_b200.utils` instead of ``util.utils`` (INTEGRATION.md shows the one-line import swap
a maintainer makes in the reference file itself; this module is that file after the swap).

The OCR pre-step (``check_ocr_box``: EasyOCR / PaddleOCR, ref:util/utils.py:514-549) is outside the accelerated path
(SURVEY.md §8f-3): it is taken from the reference if ``util.utils`` is importable, or injected with ``config['ocr_fn']``
(a callable ``image -> ((texts, xyxy_boxes), _)`` with the signature of ``check_ocr_box``).

Extra config keys (all optional): ``device``, ``detector_precision`` ("fp16" | "fp16x3"), ``caption_precision``,
``tokenizer_path``, ``allow_id_captions``, ``ocr_fn``.
"""
from __future__ import annotations

import base64
import io
import threading
from typing import Dict

from PIL import Image

from .utils import get_caption_model_processor, get_som_labeled_img, get_yolo_model


def _reference_check_ocr_box():
    """The OCR pre-step (ref:util/utils.py:514-549): :func:`omniparser_b200.ocr.check_ocr_box`, the reference's function
    with lazily bound engines (it raises a clear error if neither EasyOCR nor an injected reader is available)."""
    from .ocr import check_ocr_box
    return check_ocr_box


class Omniparser(object):
    def __init__(self, config: Dict):
        self.config = config
        device = config.get("device", "cuda")
        self.som_model = get_yolo_model(model_path=config.get("som_model_path"), device=device,
                                        precision=config.get("detector_precision"))
        self.caption_model_processor = get_caption_model_processor(
            model_name=config["caption_model_name"], model_name_or_path=config["caption_model_path"], device=device,
            precision=config.get("caption_precision", "fp16x3"), tokenizer_path=config.get("tokenizer_path"),
            allow_id_captions=config.get("allow_id_captions"))
        self._ocr = config.get("ocr_fn")
        self._lock = threading.Lock()   # ref:gradio_demo.py calls from worker threads: one parse at a time per instance
        print("Omniparser initialized!!!")

    def parse(self, image_base64: str):
        image_bytes = base64.b64decode(image_base64)
        image = Image.open(io.BytesIO(image_bytes))
        print("image size:", image.size)

        box_overlay_ratio = max(image.size) / 3200
        draw_bbox_config = {
            "text_scale": 0.8 * box_overlay_ratio,
            "text_thickness": max(int(2 * box_overlay_ratio), 1),
            "text_padding": max(int(3 * box_overlay_ratio), 1),
            "thickness": max(int(3 * box_overlay_ratio), 1),
        }
        ocr = self._ocr or _reference_check_ocr_box()
        (text, ocr_bbox), _ = ocr(image, display_img=False, output_bb_format="xyxy", easyocr_args={"text_threshold": 0.8},
                                  use_paddleocr=False)
        with self._lock:
            dino_labled_img, label_coordinates, parsed_content_list = get_som_labeled_img(
                image, self.som_model, BOX_TRESHOLD=self.config["BOX_TRESHOLD"], output_coord_in_ratio=True, ocr_bbox=ocr_bbox,
                draw_bbox_config=draw_bbox_config, caption_model_processor=self.caption_model_processor, ocr_text=text,
                use_local_semantics=True, iou_threshold=0.7, scale_img=False, batch_size=128)
        return dino_labled_img, parsed_content_list
