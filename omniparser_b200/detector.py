"""Drop-in for ``util.yolov9.YOLOv9Detector`` (ref:util/yolov9.py:27-136) running on the B200 kernels.

Same constructor / ``predict`` signature and result objects (``[Result(Boxes(xyxy, conf))]`` with device
tensors), same letterbox geometry, strict ``>`` confidence filter, un-letterbox, ``batched_nms`` semantics,
``[:max_det]`` and clamp.  Everything between the u8 image and the final boxes runs on the GPU without a host
sync: LANCZOS letterbox -> YOLOv9-E forward (CUDA graph) -> decode/filter -> bitmask NMS.
"""
from __future__ import annotations

import threading
from pathlib import Path
from typing import Dict, List, Sequence, Union

import numpy as np
import torch

from . import ops
from .yolo_engine import YoloPlan, YoloWeights, rename_upstream


class Boxes:   # ref:util/yolov9.py:16-19
    def __init__(self, xyxy: torch.Tensor, confidence: torch.Tensor):
        self.xyxy = xyxy
        self.conf = confidence


class Result:  # ref:util/yolov9.py:22-24
    def __init__(self, boxes: Boxes):
        self.boxes = boxes


def _geometry(w: int, h: int, imgsz):
    """ref:util/yolov9.py:52-61,73-80."""
    if isinstance(imgsz, int):
        tw = th = imgsz
    elif len(imgsz) == 2:
        th, tw = imgsz
    else:
        raise ValueError(f"Expected one or two image dimensions, got {imgsz}")
    tw = ((int(tw) + 31) // 32) * 32
    th = ((int(th) + 31) // 32) * 32
    scale = min(tw / w, th / h)
    rw, rh = int(w * scale), int(h * scale)
    return tw, th, scale, rw, rh, (tw - rw) // 2, (th - rh) // 2


class B200YOLOv9Detector:
    strides = (8, 16, 32)
    CAND_CAP = 16384
    OCR_CAP = 256      # OCR boxes per screenshot the device overlap filter takes (host_glue.OCR_DEVICE_CAP); more -> host path

    def __init__(self, model_path: Union[str, Path, None] = None, device: Union[str, torch.device, None] = None,
                 state_dict: Dict[str, torch.Tensor] | None = None, use_graph: bool = True, precision: str | None = None):
        """precision: "fp16" (default; the reference's own CUDA path is fp16 autocast, ref:util/yolov9.py:110-113) or
        "fp16x3" (parity grade: reproduces the fp32 CPU path's kept boxes; ~3x the tensor work).  Default from the
        environment variable B2P_DETECTOR_PRECISION."""
        import os
        self.precision = precision or os.environ.get("B2P_DETECTOR_PRECISION", "fp16")
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda" or not torch.cuda.is_available():
            # ref:util/yolov9.py:40-41 raises when CUDA is requested but unavailable; this build has no CPU path.
            raise RuntimeError(f"B200 detector needs a CUDA device (requested {self.device}); there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if state_dict is None:
            if model_path is None:
                raise FileNotFoundError("model_path (TorchScript archive or state_dict file) is required: no network here")
            self.model_path = Path(model_path)
            try:
                state_dict = torch.jit.load(str(self.model_path), map_location="cpu").state_dict()
            except Exception:
                state_dict = torch.load(str(self.model_path), map_location="cpu")
        if any(k.startswith("model.") for k in state_dict):
            state_dict = rename_upstream(state_dict)
        with torch.cuda.device(self.device):
            self.weights = YoloWeights(state_dict, self.device, precision=self.precision)
        self.model = self.weights   # attribute the reference exposes (inspected in demo.ipynb)
        self.use_graph = use_graph
        self._plans: Dict[tuple, YoloPlan] = {}
        self._io: Dict[tuple, dict] = {}
        # One parse at a time per detector handle: the reference's callers share ONE module-level model between worker
        # threads (ref:gradio_demo.py:15-16,35-59) and its PyTorch modules are re-entrant; this object's launch plans and io
        # buffers are not, so every public entry point serialises on this lock (SURVEY.md §8b "self-serialising per device").
        self._lock = threading.RLock()

    def to(self, device):   # called at ref:eval/ss_pro_gpt4o_omniv2.py:30
        return self

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _load_image(source) -> np.ndarray:
        from PIL import Image
        if isinstance(source, np.ndarray):
            return np.asarray(Image.fromarray(source).convert("RGB"))
        if isinstance(source, Image.Image):
            return np.asarray(source.convert("RGB"))
        with Image.open(source) as im:
            return np.asarray(im.convert("RGB"))

    def _get_io(self, B, H, W, imgsz, max_det, slot=0):
        tw, th, scale, rw, rh, pl, pt = _geometry(W, H, imgsz)
        key = (B, H, W, tw, th, max_det, slot)
        io = self._io.get(key)
        if io is None:
            dev = self.device
            pk = (B, th, tw)
            if pk not in self._plans:
                self._plans[pk] = YoloPlan(self.weights, B, th, tw, self.use_graph)
            plan = self._plans[pk]
            A = sum(h * w for h, w in plan.hw)
            cap = min(self.CAND_CAP, A)
            f32 = dict(dtype=torch.float32, device=dev)
            io = dict(
                plan=plan, geom=(tw, th, scale, rw, rh, pl, pt), cap=cap,
                host=torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(),
                host_count=torch.zeros((B,), dtype=torch.int32).pin_memory(),
                host_box=torch.zeros((B, max_det, 4), dtype=torch.float32).pin_memory(),
                host_cand=torch.zeros((B,), dtype=torch.int32).pin_memory(),
                src=torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
                tmp=torch.empty((B, H, max(rw, 1), 3), dtype=torch.uint8, device=dev),
                pad_l=torch.full((B,), float(pl), **f32), pad_t=torch.full((B,), float(pt), **f32),
                scale=torch.full((B,), float(np.float32(scale)), **f32),
                img_w=torch.full((B,), float(W), **f32), img_h=torch.full((B,), float(H), **f32),
                cand_box=torch.empty((B, cap, 4), **f32), cand_score=torch.empty((B, cap), **f32),
                cand_cls=torch.empty((B, cap), dtype=torch.int32, device=dev),
                cand_count=torch.zeros((B,), dtype=torch.int32, device=dev),
                keep=torch.empty((B, max_det), dtype=torch.int32, device=dev),
                out_box=torch.empty((B, max_det, 4), **f32), out_score=torch.empty((B, max_det), **f32),
                out_count=torch.zeros((B,), dtype=torch.int32, device=dev),
            )
            # device overlap filter (b2p_overlap_filter, ref:util/utils.py:241-319): OCR ratio boxes in, flags + crop list out
            mo = self.OCR_CAP
            i32 = dict(dtype=torch.int32, device=dev)
            io.update(
                max_ocr=mo,
                ocr_ratio=torch.zeros((B, mo, 4), **f32), ocr_count=torch.zeros((B,), **i32),
                host_ocr_ratio=torch.zeros((B, mo, 4), dtype=torch.float32).pin_memory(),
                host_ocr_count=torch.zeros((B,), dtype=torch.int32).pin_memory(),
                icon_state=torch.zeros((B, max_det), **i32), label_mask=torch.zeros((B, max_det, mo // 32), **i32),
                ocr_removed=torch.zeros((B, mo), **i32), icon_ratio=torch.zeros((B, max_det, 4), **f32),
                crop_box=torch.zeros((B * max_det, 4), **f32), crop_img=torch.zeros((B * max_det,), **i32),
                crop_counts=torch.zeros((B + 1,), **i32), arrive=torch.zeros((1,), **i32),
                host_state=torch.zeros((B, max_det), dtype=torch.int32).pin_memory(),
                host_mask=torch.zeros((B, max_det, mo // 32), dtype=torch.int32).pin_memory(),
                host_removed=torch.zeros((B, mo), dtype=torch.int32).pin_memory(),
                host_ratio=torch.zeros((B, max_det, 4), dtype=torch.float32).pin_memory(),
                host_crop_counts=torch.zeros((B + 1,), dtype=torch.int32).pin_memory(),
            )
            self._io[key] = io
        return io

    def detect_device(self, io, B, H, W, conf, iou, max_det):
        """Device-resident u8 images in io['src'] -> NMS outputs in io (no host sync)."""
        plan: YoloPlan = io["plan"]
        tw, th, scale, rw, rh, pl, pt = io["geom"]
        ops.letterbox(io["src"], B, H, W, rw, rh, tw, th, pl, pt, io["tmp"], plan.canvas)
        plan.run()
        ops.yolo_decode(plan.cls_out, plan.box_out, plan.hw, self.weights.nc, B, float(conf), io["pad_l"], io["pad_t"],
                        io["scale"], io["cap"], io["cand_box"], io["cand_score"], io["cand_cls"], io["cand_count"])
        ops.batched_nms(io["cand_box"], io["cand_score"], io["cand_cls"], io["cand_count"], B, io["cap"], iou, max_det,
                        io["img_w"], io["img_h"], io["keep"], io["out_box"], io["out_score"], io["out_count"])

    def filter_device(self, io, B, H, W, ocr_elems, iou_threshold, max_det=300) -> bool:
        """After :meth:`detect_device`, on the same stream: the reference's overlap filter (ref:util/utils.py:241-319,
        :411-415, :444-451) on the device -> per-icon state, OCR label masks / removed flags and the batch's crop list, plus
        the async D2H of the flags into this io slot's pinned mirrors.  ``ocr_elems[i]`` = host_glue.ocr_elements(...) of
        screenshot i (the strings stay on the host).  Returns False -- nothing launched -- when a screenshot has more OCR boxes
        than the kernel takes; the caller then runs the host list logic (host_glue.build_elements) instead."""
        mo = io["max_ocr"]
        if any(len(e) > mo for e in ocr_elems):
            return False
        hr, hc = io["host_ocr_ratio"], io["host_ocr_count"]
        for i, e in enumerate(ocr_elems):
            hc[i] = len(e)
            if e:
                hr[i, :len(e)] = torch.tensor([x["bbox"] for x in e], dtype=torch.float32)
        io["ocr_ratio"].copy_(hr, non_blocking=True)
        io["ocr_count"].copy_(hc, non_blocking=True)
        ops.overlap_filter(io["out_box"], io["out_count"], B, max_det, io["img_w"], io["img_h"], io["ocr_ratio"], io["ocr_count"],
                           mo, float(iou_threshold), io["icon_state"], io["label_mask"], io["ocr_removed"], io["icon_ratio"],
                           io["crop_box"], io["crop_img"], io["crop_counts"], io["arrive"])
        for h_, d_ in (("host_state", "icon_state"), ("host_mask", "label_mask"), ("host_removed", "ocr_removed"),
                       ("host_ratio", "icon_ratio"), ("host_crop_counts", "crop_counts")):
            io[h_].copy_(io[d_], non_blocking=True)
        return True

    @staticmethod
    def elements_from_io(io, i, n_det, ocr_elem):
        """Screenshot i of a filtered batch (after the D2H above has completed) -> the reference's sorted element list."""
        from . import host_glue
        return host_glue.elements_from_flags(io["host_ratio"][i, :n_det].tolist(), io["host_state"][i, :n_det].tolist(),
                                             (io["host_mask"][i, :n_det].numpy().view("uint32")), ocr_elem,
                                             io["host_removed"][i, :len(ocr_elem)].tolist())

    @staticmethod
    def check_capacity(cand_count_host: torch.Tensor, cap: int) -> None:
        """The decode kernel keeps the first ``cap`` candidates in anchor order; more than that (only reachable with
        full-resolution ``imgsz`` and a very low threshold) would silently drop the coarse-stride candidates."""
        m = int(cand_count_host.max()) if cand_count_host.numel() else 0
        if m > cap:
            raise RuntimeError(f"{m} candidates above the confidence threshold exceed the NMS capacity ({cap}): raise the "
                               "threshold or lower imgsz")

    @torch.inference_mode()
    def predict_batch(self, images: Sequence[np.ndarray], conf=0.25, imgsz=640, iou=0.7, max_det=300) -> List[Result]:
        """Same-size u8 HWC images -> one Result per image (one H2D copy, one D2H of the counts)."""
        B = len(images)
        H, W = images[0].shape[:2]
        with self._lock, torch.cuda.device(self.device):
            io = self._get_io(B, H, W, imgsz, max_det)
            for i, im in enumerate(images):
                assert im.shape == (H, W, 3) and im.dtype == np.uint8
                io["host"][i].copy_(torch.from_numpy(np.ascontiguousarray(im)))
            io["src"].copy_(io["host"], non_blocking=True)
            self.detect_device(io, B, H, W, conf, iou, max_det)
            counts = io["out_count"].cpu().tolist()
            self.check_capacity(io["cand_count"].cpu(), io["cap"])
            return [Result(Boxes(io["out_box"][i, :n].clone(), io["out_score"][i, :n].clone())) for i, n in enumerate(counts)]

    @torch.inference_mode()
    def predict(self, source, conf=0.25, imgsz=640, iou=0.7, max_det=300):
        """ref:util/yolov9.py:115-136."""
        return self.predict_batch([self._load_image(source)], conf, imgsz, iou, max_det)
