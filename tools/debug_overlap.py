"""Debug aid: reproduce tests/test_pipeline_gpu.py::test_device_overlap_filter_path_equals_host_list_logic and dump the detector
boxes, OCR lists and the device filter's flags for offline comparison with oracle.ref_restate.overlap_flags_loops."""
import json, sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
from omniparser_b200 import synth, host_glue
from omniparser_b200 import utils as U
det, cmp_ = ge.standin_models(torch.device("cuda", 0))
seeds = [60, 61, 62, 63]
imgs = [synth.screenshot(s) for s in seeds]
ocr = []
probe = det.predict_batch(imgs, conf=0.05, iou=0.1)
for s, r in zip(seeds, probe):
    texts, boxes = synth.ocr_boxes(s)
    for j, b in enumerate(r.boxes.xyxy.cpu().tolist()[:12]):
        x1, y1, x2, y2 = b
        if j % 3 == 0 and x2 - x1 > 12 and y2 - y1 > 12:
            boxes.append([int(x1) + 3, int(y1) + 3, int(x2) - 3, int(y2) - 3]); texts.append(f"in{j}")
        elif j % 3 == 1:
            boxes.append([max(0, int(x1) - 20), max(0, int(y1) - 20), int(x2) + 20, int(y2) + 20]); texts.append(f"around{j}")
    ocr.append((texts, boxes))
U._HOST_GLUE = False
res = U.parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
torch.cuda.synchronize()
io_ = det._get_io(4, 1080, 1920, 640, 300)
dump = dict(ocr=ocr, counts=io_["out_count"].cpu().tolist(), boxes=io_["out_box"].cpu().tolist(), state=io_["icon_state"].cpu().tolist(),
            mask=io_["label_mask"].cpu().tolist(), removed=io_["ocr_removed"].cpu().tolist(), ratio=io_["icon_ratio"].cpu().tolist(),
            crop_counts=io_["crop_counts"].cpu().tolist(), elems_device=[el for el, _ in res])
U._HOST_GLUE = True
res2 = U.parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
dump["elems_host"] = [el for el, _ in res2]
json.dump(dump, open("gpurun_out/debug_overlap.json", "w"))
for i in range(4):
    a = [(e["bbox"], e["source"]) for e in dump["elems_device"][i]]; b = [(e["bbox"], e["source"]) for e in dump["elems_host"][i]]
    print("shot", i, "device", len(a), "host", len(b), "equal", a == b)
