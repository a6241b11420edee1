"""One parse step (batch of 8 screenshots) between cudaProfilerStart/Stop, eager launches (B2P_NO_GRAPH=1) so that
the B2P_DEBUG gemm shape log lines map 1:1 onto the ncu launch list.
  B2P_NO_GRAPH=1 B2P_DEBUG=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/step_launches.csv python tools/profile_step.py 2> gpurun_out/step_shapes.log
"""
import os, sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
from omniparser_b200 import synth
from omniparser_b200.utils import parse_screenshots, ParseTimings

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
dev = torch.device("cuda", 0)
det, cmp_ = ge.standin_models(dev, prec)
B = 8
imgs = [synth.screenshot(s) for s in range(B)]
ocr = [synth.ocr_boxes(s) for s in range(B)]
for _ in range(2):
    parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8)
torch.cuda.synchronize()
print("PROFILE_START", file=sys.stderr, flush=True)
torch.cuda.profiler.start()
tm = ParseTimings()
parse_screenshots(imgs, det, cmp_, ocr, BOX_TRESHOLD=0.05, iou_threshold=0.7, max_new_tokens=8, timings=tm)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("PROFILE_STOP", dict(tm), file=sys.stderr, flush=True)
