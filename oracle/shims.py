"""ORACLE (test infrastructure): import the UNMODIFIED reference glue in this container.

``/root/reference/util/utils.py`` imports four non-arithmetic packages that are
not installed (``easyocr``, ``paddleocr``, ``matplotlib``, ``supervision``) and
instantiates OCR engines at import time (ref:util/utils.py:20-31).  This module
injects minimal stand-ins into ``sys.modules`` so the reference's own
``get_som_labeled_img`` / ``YOLOv9Detector`` / ``get_parsed_content_icon`` run
unmodified as the oracle.  Only available where ``/root/reference`` exists (this
container); the GPU box uses the restatements in ``oracle/ref_restate.py`` and
the golden fixtures generated here (``oracle/make_golden.py``).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE / "util" / "utils.py").is_file()


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install_shims() -> None:
    if "supervision" in sys.modules and getattr(sys.modules["supervision"], "_b2p_shim", False):
        return
    import numpy as np

    easyocr = _mod("easyocr")

    class Reader:  # constructed at import, ref:util/utils.py:22
        def __init__(self, *a, **k):
            pass

        def readtext(self, image, **kw):
            return []

    easyocr.Reader = Reader

    paddleocr = _mod("paddleocr")

    class PaddleOCR:  # constructed at import, ref:util/utils.py:23-31
        def __init__(self, *a, **k):
            pass

        def ocr(self, image, cls=False):
            return [[]]

    paddleocr.PaddleOCR = PaddleOCR

    mpl = _mod("matplotlib")
    plt = _mod("matplotlib.pyplot")
    mpl.pyplot = plt

    openai = sys.modules.get("openai")
    if openai is None:
        try:
            import openai  # noqa: F401
        except Exception:
            openai = _mod("openai")
            openai.AzureOpenAI = object

    sv = _mod("supervision")
    sv._b2p_shim = True
    core = _mod("supervision.detection")
    core2 = _mod("supervision.detection.core")
    draw = _mod("supervision.draw")
    color = _mod("supervision.draw.color")

    class Detections:
        def __init__(self, xyxy, class_id=None, **kw):
            self.xyxy = np.asarray(xyxy)
            self.class_id = class_id
            self.confidence = None
            self.tracker_id = None

        def __len__(self):
            return len(self.xyxy)

    class Color:
        def __init__(self, r=0, g=0, b=0):
            self.r, self.g, self.b = r, g, b

        def as_bgr(self):
            return (self.b, self.g, self.r)

        def as_rgb(self):
            return (self.r, self.g, self.b)

        @classmethod
        def black(cls):
            return cls(0, 0, 0)

        @classmethod
        def white(cls):
            return cls(255, 255, 255)

    class ColorPalette:
        def __init__(self, colors):
            self.colors = colors

        def by_idx(self, i):
            return self.colors[i % len(self.colors)]

    Color.BLACK = Color(0, 0, 0)
    Color.WHITE = Color(255, 255, 255)
    # supervision 0.18.0 (pinned at ref:requirements.txt) ColorPalette.DEFAULT, recalled (DEFAULT_COLOR_PALETTE hex list)
    _hex = ["A351FB", "FF4040", "FFA1A0", "FF7633", "FFB633", "D1D435", "4CFB12", "94CF1A", "40DE8A", "1B9640", "00D6C1",
            "2E9CAA", "00C4FF", "364797", "6675FF", "0019EF", "863AFF", "530087", "CD3AFF", "FF97CA", "FF39C9"]
    ColorPalette.DEFAULT = ColorPalette([Color(int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16)) for h in _hex])
    sv.Detections = Detections
    core2.Detections = Detections
    sv.detection = core
    core.core = core2
    sv.draw = draw
    draw.color = color
    color.Color = Color
    color.ColorPalette = ColorPalette
    sv.Color = Color
    sv.ColorPalette = ColorPalette


def import_reference():
    """Returns (util.utils, util.yolov9) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    install_shims()
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    import util.utils as ref_utils  # noqa: E402
    import util.yolov9 as ref_yolov9  # noqa: E402

    return ref_utils, ref_yolov9
