"""Debug: pipelined vs sequential parity per batch for 1 and 2 caption lanes (prints which batches / rows differ)."""
import sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
from omniparser_b200 import synth
from omniparser_b200.utils import PipelinedParser, parse_screenshots

DEV = torch.device("cuda", 0)
det, cmp_ = ge.standin_models(DEV)
batches = []
for b in range(6):
    seeds = [40 + 2 * b, 41 + 2 * b]
    batches.append(([synth.screenshot(s) for s in seeds], [synth.ocr_boxes(s) for s in seeds]))
ref = [parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8) for imgs, ocr in batches]
ref2 = [parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8) for imgs, ocr in batches]
for bi, (a, b) in enumerate(zip(ref, ref2)):
    for (e1, i1), (e2, i2) in zip(a, b):
        if not torch.equal(i1, i2):
            print("SEQUENTIAL path not reproducible at batch", bi)
for lanes in [int(a) for a in sys.argv[1:]] or [1, 2]:
    pp = PipelinedParser(det, cmp_, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8, caption_lanes=lanes)
    for rep in range(2):
        got = list(pp.run(iter(batches)))
        for bi, (gb, rb) in enumerate(zip(got, ref)):
            for si, ((ge_, gi), (re_, ri)) in enumerate(zip(gb, rb)):
                same_boxes = [e["bbox"] for e in ge_] == [e["bbox"] for e in re_]
                if gi.shape != ri.shape or not torch.equal(gi, ri):
                    bad = (gi[:, :min(gi.shape[1], ri.shape[1])] != ri[:, :min(gi.shape[1], ri.shape[1])]).any(1).nonzero().flatten().tolist() if gi.shape[0] == ri.shape[0] else "row count differs"
                    print(f"lanes={lanes} rep={rep} batch={bi} (lane {bi % lanes}) shot={si}: ids differ, shapes {tuple(gi.shape)} vs {tuple(ri.shape)}, same_boxes={same_boxes}, bad rows {bad if isinstance(bad, str) else (len(bad), bad[:8])}")
        print(f"lanes={lanes} rep={rep} done", flush=True)
