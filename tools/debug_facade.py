"""Debug aid: run the facade-golden case (seed 7) through the parity-grade detector and dump detector output + parsed list."""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from PIL import Image
from omniparser_b200 import synth
from omniparser_b200.detector import B200YOLOv9Detector
from standin.yolo_weights import yolo_standin
g = json.load(open("tests/golden/facade_seed7.json"))
w, h = g["case"]["size"]
img = synth.screenshot(g["case"]["seed"], w, h)
out = {}
for prec in ("fp16x3", "fp16"):
    det = B200YOLOv9Detector(state_dict=yolo_standin(0).state_dict(), device="cuda:0", precision=prec)
    for conf in (0.05, 0.03):
        b = det.predict(Image.fromarray(img), conf=conf, iou=0.1)[0].boxes
        out[f"{prec}_{conf}"] = dict(xyxy=b.xyxy.cpu().tolist(), conf=b.conf.cpu().tolist())
json.dump(out, open("gpurun_out/debug_facade.json", "w"))
print("dumped", {k: len(v["conf"]) for k, v in out.items()})
