"""Seeded stand-in weights for the detector (test/benchmark data, not product; see standin/__init__.py).

Real ``icon_detect_v3/model.pt`` and ``icon_caption_florence`` weights are not
in this environment (no network); parity is therefore pinned on seeded
stand-ins with the exact layer shapes.  To keep a deep random network
well-conditioned (spatial signal alive at every depth, activations O(1) in
fp16) the BatchNorm running statistics are calibrated once on seeded synthetic
screenshots and committed as ``tests/golden/yolov9e_bn_calib_seed0.npz`` so
every machine rebuilds bit-identical weights (the conv weights come from
``torch.Generator`` on CPU, which is deterministic; data-dependent statistics
are not, hence the fixture).  Script that made the fixture: this file,
``python -m standin.yolo_weights``.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from .yolov9e import YOLOv9E

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
BN_CALIB = GOLDEN / "yolov9e_bn_calib_seed0.npz"

# Head calibration constants of the stand-in (chosen once so that the synthetic
# 1920x1080 screenshots give ~60 post-NMS boxes at BOX_TRESHOLD=0.05, iou=0.1).
DFL_BIN_SLOPE = -0.5          # final box-conv bias = slope * bin  -> ~1.5 strides per side
CLS_BIAS = (-5.6, -8.5, -10.5)  # per-scale class-logit bias (stride 8, 16, 32)
CLS_GAIN = 3.0


def _seed_weights(m: YOLOv9E, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels // mod.groups * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                if mod.bias is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            elif isinstance(mod, nn.BatchNorm2d):
                n = mod.num_features
                # gamma ~0.5, beta ~0.5 keeps SiLU in a mildly non-linear regime: a random deep net with unit-gain
                # BN is chaotic (measured: fp16 rounding noise amplified ~1000x over the 100-layer depth, 50 % drift at
                # the heads), which says nothing about kernel correctness; trained detectors are not in that regime.
                mod.weight.copy_(0.5 + 0.05 * torch.randn(n, generator=g))
                mod.bias.copy_(0.5 + 0.2 * torch.randn(n, generator=g))
        for i, seq in enumerate(m.detect.cv2):
            bins = torch.arange(16, dtype=torch.float32).repeat(4)
            seq[-1].bias.copy_(DFL_BIN_SLOPE * bins)
        for i, seq in enumerate(m.detect.cv3):
            seq[-1].weight.mul_(CLS_GAIN)
            seq[-1].bias.fill_(CLS_BIAS[i])


def calib_images(n: int = 2) -> torch.Tensor:
    """Letterboxed 640x640 synthetic screenshots, f32 NCHW in [0,1] (box resample; calibration only)."""
    from PIL import Image
    from omniparser_b200 import synth

    out = []
    for s in range(n):
        im = Image.fromarray(synth.screenshot(1000 + s)).resize((640, 360), Image.Resampling.BOX)
        canvas = Image.new("RGB", (640, 640), (114, 114, 114))
        canvas.paste(im, (0, 140))
        out.append(torch.from_numpy(np.asarray(canvas, dtype=np.float32).transpose(2, 0, 1) / 255.0))
    return torch.stack(out)


def _calibrate_bn(m: YOLOv9E) -> dict:
    bns = [b for b in m.modules() if isinstance(b, nn.BatchNorm2d)]
    for b in bns:
        b.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(calib_images())
    m.eval()
    stats = {}
    for i, b in enumerate(bns):
        stats[f"m{i}"] = b.running_mean.numpy().copy()
        stats[f"v{i}"] = b.running_var.numpy().copy()
    return stats


def yolo_standin(seed: int = 0, nc: int = 1) -> YOLOv9E:
    m = YOLOv9E(nc).eval()
    _seed_weights(m, seed)
    if seed == 0 and BN_CALIB.is_file():
        stats = np.load(BN_CALIB)
    else:
        if seed == 0:
            raise FileNotFoundError(f"{BN_CALIB} missing; run `python -m standin.yolo_weights` where it can be regenerated")
        stats = _calibrate_bn(m)
    bns = [b for b in m.modules() if isinstance(b, nn.BatchNorm2d)]
    with torch.no_grad():
        for i, b in enumerate(bns):
            b.running_mean.copy_(torch.from_numpy(stats[f"m{i}"]))
            b.running_var.copy_(torch.from_numpy(stats[f"v{i}"]))
    return m.eval()


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    m = YOLOv9E(1).eval()
    _seed_weights(m, 0)
    stats = _calibrate_bn(m)
    np.savez_compressed(BN_CALIB, **stats)
    print("wrote", BN_CALIB, os.path.getsize(BN_CALIB))


if __name__ == "__main__":
    main()
