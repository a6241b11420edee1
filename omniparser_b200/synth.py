"""Seeded synthetic 1920x1080 "UI-like" screenshots + OCR boxes.

The reference ships no benchmark inputs; BASELINE.json's metric is quoted on
synthetic 1920x1080 screenshots with ~60 boxes each.  ``screenshot(seed)``
draws a flat/gradient background, ~60 icon-like rounded blobs of side
U[20,80] px on a jittered grid and a few text-like bars, all from
``numpy.random.default_rng(seed)`` so every process (and the CPU baseline)
sees identical bytes.  ``ocr_boxes(seed)`` returns the 20 synthetic OCR boxes
the reference needs as *input* (it crashes with none, ref:util/utils.py:437-444).
"""
from __future__ import annotations

import numpy as np

W, H = 1920, 1080


def screenshot(seed: int, w: int = W, h: int = H, n_icons: int = 60) -> np.ndarray:
    rng = np.random.default_rng(seed)
    base = rng.integers(180, 250, size=3)
    yy = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None, None]
    xx = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :, None]
    tilt = rng.uniform(-25, 25, size=(2, 3)).astype(np.float32)
    img = base[None, None, :].astype(np.float32) + yy * tilt[0] + xx * tilt[1]
    img = np.broadcast_to(img, (h, w, 3)).copy()
    # title bar / side bar
    img[: h // 24] = rng.integers(30, 90, size=3)
    img[:, : w // 40] = rng.integers(60, 120, size=3)
    # icons on a jittered grid
    cols = int(np.ceil(np.sqrt(n_icons * w / h)))
    rows = int(np.ceil(n_icons / cols))
    cw, ch = (w - 120) / cols, (h - 120) / rows
    k = 0
    for r in range(rows):
        for c in range(cols):
            if k >= n_icons:
                break
            k += 1
            sw, sh = rng.integers(20, 81, size=2)
            cx = 60 + (c + 0.5) * cw + rng.uniform(-0.15, 0.15) * cw
            cy = 60 + (r + 0.5) * ch + rng.uniform(-0.15, 0.15) * ch
            x0, y0 = int(cx - sw / 2), int(cy - sh / 2)
            x1, y1 = min(x0 + sw, w), min(y0 + sh, h)
            x0, y0 = max(x0, 0), max(y0, 0)
            col = rng.integers(0, 256, size=3).astype(np.float32)
            img[y0:y1, x0:x1] = col
            # glyph: a contrasting inner blob and a diagonal stroke
            ix0, iy0 = x0 + (x1 - x0) // 4, y0 + (y1 - y0) // 4
            ix1, iy1 = x1 - (x1 - x0) // 4, y1 - (y1 - y0) // 4
            img[iy0:iy1, ix0:ix1] = 255.0 - col
            n = min(ix1 - ix0, iy1 - iy0)
            for t in range(max(n, 0)):
                img[iy0 + t, ix0 + t] = col
    # text-like bars
    for _ in range(24):
        bx, by = rng.integers(40, w - 260), rng.integers(40, h - 40)
        bw, bh = rng.integers(60, 220), rng.integers(8, 18)
        img[by : by + bh, bx : bx + bw : 3] = rng.integers(0, 80)
    noise = rng.integers(-3, 4, size=(h, w, 1))
    return np.clip(img + noise, 0, 255).astype(np.uint8)


def ocr_boxes(seed: int, w: int = W, h: int = H, n: int = 20):
    """(texts, xyxy pixel boxes) in the format ``check_ocr_box`` returns (ref:util/utils.py:514-549)."""
    rng = np.random.default_rng(10_000 + seed)
    boxes, texts = [], []
    for i in range(n):
        x = int(rng.integers(0, w - 70))
        y = int(rng.integers(0, h - 30))
        boxes.append([x, y, x + 60, y + 20])
        texts.append(f"t{i}")
    return texts, boxes
