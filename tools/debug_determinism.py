"""Debug aid: is the detector bit-deterministic run to run, and do parse_screenshots (host list logic) / (device filter) see the
same detector output?"""
import sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
from omniparser_b200 import synth
from omniparser_b200 import utils as U
det, cmp_ = ge.standin_models(torch.device("cuda", 0), "fp16x3")
det = ge.standin_models(torch.device("cuda", 0))[0]
seeds = [60, 61, 62, 63]
imgs = [synth.screenshot(s) for s in seeds]
ocr = [synth.ocr_boxes(s) for s in seeds]
outs = []
for rep in range(4):
    r = det.predict_batch(imgs, conf=0.05, iou=0.1)
    outs.append([x.boxes.xyxy.clone().cpu() for x in r])
for rep in range(1, 4):
    same = all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(outs[0], outs[rep]))
    print("predict_batch rep", rep, "bit-identical to rep 0:", same)
io_ = det._get_io(4, 1080, 1920, 640, 300)
snaps = []
for mode in (True, False, True, False):
    U._HOST_GLUE = mode
    res = U.parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
    torch.cuda.synchronize()
    snaps.append((io_["out_count"].cpu().clone(), io_["out_box"].cpu().clone(), [[e["bbox"] for e in el] for el, _ in res]))
    print("host_glue" if mode else "device   ", "counts", snaps[-1][0].tolist(), "n_elems", [len(x) for x in snaps[-1][2]])
for i in range(1, 4):
    c0, b0, e0 = snaps[0]; c, b, e = snaps[i]
    eq = torch.equal(c0, c) and all(torch.equal(b0[k, :c0[k]], b[k, :c[k]]) for k in range(4) if c0[k] == c[k])
    print("run", i, "detector output identical to run 0:", eq, "elements identical:", e == e0)
r = det.predict_batch(imgs, conf=0.05, iou=0.1)
print("predict_batch after parse identical to first:", all(torch.equal(a, x.boxes.xyxy.cpu()) for a, x in zip(outs[0], r)))
print("predict vs parse run0 boxes identical:", all(torch.equal(outs[0][k], snaps[0][1][k, :snaps[0][0][k]]) for k in range(4)))
