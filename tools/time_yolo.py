"""Times the YOLOv9-E plan (CUDA events, graph replay) for a few batch sizes. Usage: python tools/time_yolo.py [B ...]"""
import sys, time
import torch
sys.path.insert(0, ".")
from omniparser_b200.yolo_engine import YoloPlan, YoloWeights
from standin.yolo_weights import yolo_standin

dev = torch.device("cuda:0")
w = YoloWeights(yolo_standin(0).state_dict(), dev)
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    for graph in (False, True):
        plan = YoloPlan(w, B, 640, 640, use_graph=graph)
        plan.canvas.random_(0, 255)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B} graph={graph} launches={plan.n_launches} {ms:.3f} ms/forward  {plan.flops / ms / 1e9:.1f} TFLOP/s "
              f"(host wall {1e3 * (time.perf_counter() - t0) / n:.3f} ms)", flush=True)
