"""PyTorch fp32 restatement of YOLOv9-E: the CPU oracle's detector network AND the definition of the seeded stand-in
checkpoint (test/benchmark infrastructure, not product; see standin/__init__.py).

The reference runs its detector as an opaque TorchScript archive
(``icon_detect_v3/model.pt``, ref:util/yolov9.py:12-13,50) that is not in the
reference tree and not in this container.  "parity unpinned": there is no
golden vector for the network arithmetic in the reference; what the reference
does pin is the *contract* of that archive, consumed by
``YOLOv9Detector._decode`` (ref:util/yolov9.py:89-108): a tuple of six tensors
``[cls8 (B,nc,H/8,W/8), ltrb8 (B,4,H/8,W/8), cls16, ltrb16, cls32, ltrb32]``
where even entries are class logits and odd entries are LTRB distances in
stride units.

This file restates the published YOLOv9-E topology (WongKinYiu/yolov9
``models/detect/yolov9-e.yaml``, MIT lineage named in ref:README.md:71; block
definitions in that repo's ``models/common.py``) so that a seeded stand-in with
the exact layer shapes (57.3 M parameters, 188.6 GFLOP @640x640, nc=1) can be
traced to TorchScript and loaded by the unmodified reference wrapper.  The
B200 engine consumes the same ``state_dict`` (see
``omniparser_b200/yolo_plan.py``) and is compared against this module.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-3  # yolov9 initialize_weights() sets eps=1e-3 on every BatchNorm2d


def autopad(k: int, p=None) -> int:
    return k // 2 if p is None else p


class Conv(nn.Module):
    """conv(k,s,p=k//2,bias=False) + BN + SiLU (common.py::Conv)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=BN_EPS)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class RepConvN(nn.Module):
    """3x3 conv+BN and 1x1 conv+BN summed, then SiLU (common.py::RepConvN, no identity branch)."""

    def __init__(self, c1, c2, k=3, s=1, p=1):
        super().__init__()
        self.conv1 = Conv(c1, c2, k, s, p=p, act=False)
        self.conv2 = Conv(c1, c2, 1, s, p=(p - k // 2), act=False)
        self.act = nn.SiLU()

    def forward(self, x):
        return self.act(self.conv1(x) + self.conv2(x))


class RepNBottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = RepConvN(c1, c_, 3, 1)
        self.cv2 = Conv(c_, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class RepNCSP(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(RepNBottleneck(c_, c_, True, e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class RepNCSPELAN4(nn.Module):
    def __init__(self, c1, c2, c3, c4, n=1):
        super().__init__()
        self.c = c3 // 2
        self.cv1 = Conv(c1, c3, 1, 1)
        self.cv2 = nn.Sequential(RepNCSP(c3 // 2, c4, n), Conv(c4, c4, 3, 1))
        self.cv3 = nn.Sequential(RepNCSP(c4, c4, n), Conv(c4, c4, 3, 1))
        self.cv4 = Conv(c3 + 2 * c4, c2, 1, 1)

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in (self.cv2, self.cv3))
        return self.cv4(torch.cat(y, 1))


class ADown(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1 // 2, self.c, 3, 2, 1)
        self.cv2 = Conv(c1 // 2, self.c, 1, 1, 0)

    def forward(self, x):
        x = F.avg_pool2d(x, 2, 1, 0, False, True)
        x1, x2 = x.chunk(2, 1)
        x1 = self.cv1(x1)
        x2 = F.max_pool2d(x2, 3, 2, 1)
        x2 = self.cv2(x2)
        return torch.cat((x1, x2), 1)


class SPPELAN(nn.Module):
    def __init__(self, c1, c2, c3):
        super().__init__()
        self.cv1 = Conv(c1, c3, 1, 1)
        self.pool = nn.MaxPool2d(kernel_size=5, stride=1, padding=2)
        self.cv5 = Conv(4 * c3, c2, 1, 1)

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(self.pool(y[-1]))
        return self.cv5(torch.cat(y, 1))


class CBLinear(nn.Module):
    def __init__(self, c1, c2s: Sequence[int]):
        super().__init__()
        self.c2s = list(c2s)
        self.conv = nn.Conv2d(c1, sum(c2s), 1, 1, 0, bias=True)

    def forward(self, x) -> List[torch.Tensor]:
        return list(self.conv(x).split(self.c2s, dim=1))


class CBFuse(nn.Module):
    def __init__(self, idx: Sequence[int]):
        super().__init__()
        self.idx = list(idx)

    def forward(self, xs):
        target = xs[-1].shape[2:]
        res = [F.interpolate(x[self.idx[i]], size=target, mode="nearest") for i, x in enumerate(xs[:-1])]
        return torch.sum(torch.stack(res + [xs[-1]]), dim=0)


class DDetect(nn.Module):
    """Dual-branch anchor-free head; forward returns the TorchScript archive's 6-tensor contract."""

    reg_max = 16

    def __init__(self, nc=1, ch=(256, 512, 512)):
        super().__init__()
        self.nc = nc
        c2 = max(ch[0] // 4, self.reg_max * 4, 16)
        c2 = int(math.ceil(c2 / 4) * 4)
        c3 = max(ch[0], min(nc * 2, 128))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3, g=4), nn.Conv2d(c2, 4 * self.reg_max, 1, groups=4))
            for x in ch
        )
        self.cv3 = nn.ModuleList(
            nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch
        )
        self.register_buffer("proj", torch.arange(self.reg_max, dtype=torch.float32), persistent=False)

    def forward(self, feats: List[torch.Tensor]):
        outs = []
        for i, x in enumerate(feats):
            box = self.cv2[i](x)  # (B, 64, H, W): channel = side*16 + bin
            cls = self.cv3[i](x)  # (B, nc, H, W) logits
            b, _, h, w = box.shape
            dist = box.view(b, 4, self.reg_max, h, w).softmax(2)
            ltrb = (dist * self.proj.view(1, 1, -1, 1, 1)).sum(2)  # (B,4,H,W), stride units
            outs += [cls, ltrb]
        return tuple(outs)


class YOLOv9E(nn.Module):
    """Layer numbering follows yolov9-e.yaml (0 = Silence/input)."""

    def __init__(self, nc: int = 1):
        super().__init__()
        self.nc = nc
        E = RepNCSPELAN4
        self.l1 = Conv(3, 64, 3, 2)
        self.l2 = Conv(64, 128, 3, 2)
        self.l3 = E(128, 256, 128, 64, 2)
        self.l4 = ADown(256, 256)
        self.l5 = E(256, 512, 256, 128, 2)
        self.l6 = ADown(512, 512)
        self.l7 = E(512, 1024, 512, 256, 2)
        self.l8 = ADown(1024, 1024)
        self.l9 = E(1024, 1024, 512, 256, 2)
        self.l10 = CBLinear(64, [64])
        self.l11 = CBLinear(256, [64, 128])
        self.l12 = CBLinear(512, [64, 128, 256])
        self.l13 = CBLinear(1024, [64, 128, 256, 512])
        self.l14 = CBLinear(1024, [64, 128, 256, 512, 1024])
        self.l15 = Conv(3, 64, 3, 2)
        self.l16 = CBFuse([0, 0, 0, 0, 0])
        self.l17 = Conv(64, 128, 3, 2)
        self.l18 = CBFuse([1, 1, 1, 1])
        self.l19 = E(128, 256, 128, 64, 2)
        self.l20 = ADown(256, 256)
        self.l21 = CBFuse([2, 2, 2])
        self.l22 = E(256, 512, 256, 128, 2)
        self.l23 = ADown(512, 512)
        self.l24 = CBFuse([3, 3])
        self.l25 = E(512, 1024, 512, 256, 2)
        self.l26 = ADown(1024, 1024)
        self.l27 = CBFuse([4])
        self.l28 = E(1024, 1024, 512, 256, 2)
        self.l29 = SPPELAN(1024, 512, 256)
        self.l32 = E(1536, 512, 512, 256, 2)
        self.l35 = E(1024, 256, 256, 128, 2)
        self.l36 = ADown(256, 256)
        self.l38 = E(768, 512, 512, 256, 2)
        self.l39 = ADown(512, 512)
        self.l41 = E(1024, 512, 1024, 512, 2)
        self.detect = DDetect(nc, (256, 512, 512))

    def forward(self, x):
        x1 = self.l1(x)
        x2 = self.l2(x1)
        x3 = self.l3(x2)
        x5 = self.l5(self.l4(x3))
        x7 = self.l7(self.l6(x5))
        x9 = self.l9(self.l8(x7))
        r10, r11, r12, r13, r14 = self.l10(x1), self.l11(x3), self.l12(x5), self.l13(x7), self.l14(x9)
        x16 = self.l16([r10, r11, r12, r13, r14, self.l15(x)])
        x18 = self.l18([r11, r12, r13, r14, self.l17(x16)])
        x19 = self.l19(x18)
        x21 = self.l21([r12, r13, r14, self.l20(x19)])
        x22 = self.l22(x21)
        x24 = self.l24([r13, r14, self.l23(x22)])
        x25 = self.l25(x24)
        x27 = self.l27([r14, self.l26(x25)])
        x28 = self.l28(x27)
        x29 = self.l29(x28)
        x32 = self.l32(torch.cat((F.interpolate(x29, scale_factor=2.0, mode="nearest"), x25), 1))
        x35 = self.l35(torch.cat((F.interpolate(x32, scale_factor=2.0, mode="nearest"), x22), 1))
        x38 = self.l38(torch.cat((self.l36(x35), x32), 1))
        x41 = self.l41(torch.cat((self.l39(x38), x29), 1))
        return self.detect([x35, x38, x41])


def seeded_yolov9e(seed: int = 0, nc: int = 1, cls_bias: float = 0.0) -> YOLOv9E:
    """Deterministic stand-in weights with O(1) activations at every depth.

    Conv weights ~ N(0, gain/fan_in); BN gamma/beta/mean/var drawn in a mild
    range so that folding them is a non-trivial transform (tests the engine's
    BN folding) while keeping the forward well-conditioned in fp16.
    """
    g = torch.Generator().manual_seed(seed)
    m = YOLOv9E(nc).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels // mod.groups * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                if mod.bias is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            elif isinstance(mod, nn.BatchNorm2d):
                n = mod.num_features
                mod.weight.copy_(1.0 + 0.1 * torch.randn(n, generator=g))
                mod.bias.copy_(0.1 * torch.randn(n, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                mod.running_var.copy_(1.0 + 0.2 * torch.rand(n, generator=g))
        for seq in m.detect.cv3:
            seq[-1].bias.fill_(cls_bias)
    return m


def export_torchscript(model: YOLOv9E, path, example_hw=(64, 64)) -> None:
    """Write the stand-in as the TorchScript archive the reference loads (ref:util/yolov9.py:50)."""
    import os

    os.makedirs(os.path.dirname(str(path)), exist_ok=True)
    ex = torch.zeros(1, 3, *example_hw)
    with torch.no_grad():
        ts = torch.jit.trace(model, ex, check_trace=False)
    ts.save(str(path))


class UpstreamNamedYOLOv9E(nn.Module):
    """The same graph with its parameters registered under the names of an upstream WongKinYiu/yolov9 archive
    (``model.N.*``, detect head = ``model.42``): the fixture for the loader path the real
    ``weights/icon_detect_v3/model.pt`` takes (ref:util/yolov9.py:50 ``torch.jit.load``)."""

    def __init__(self, m: YOLOv9E):
        super().__init__()
        layers = [nn.Identity() for _ in range(43)]
        for n in range(1, 42):
            if hasattr(m, f"l{n}"):
                layers[n] = getattr(m, f"l{n}")
        layers[42] = m.detect
        self.model = nn.ModuleList(layers)

    def __getattr__(self, name):   # l{N} / detect resolve to model[N], so YOLOv9E.forward runs unchanged
        if name != "model":
            if name.startswith("l") and name[1:].isdigit():
                return self.model[int(name[1:])]
            if name == "detect":
                return self.model[42]
        return super().__getattr__(name)

    forward = YOLOv9E.forward
