"""Representative YOLOv9-E GEMM/conv shapes at batch 8 for `ncu --set full` (profiles/): 
  ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -o gpurun_out/prof_gemm python tools/prof_gemm.py"""
import sys
sys.path.insert(0, ".")
import torch
from omniparser_b200 import ops

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g) * 0.3).half().to(dev)
cases = []
# (name, callable)
x = ops.Map(rnd(8, 80, 80, 64), 0, 64); w = rnd(64, 576); o = ops.new_map(8, 80, 80, 64, dev); b = torch.zeros(64, device=dev)
cases.append(("conv3x3 s1 8x80x80 64->64", lambda: ops.conv3x3(x, w, o, 1, b, None, ops.ACT_SILU)))
x2 = ops.Map(rnd(8, 40, 40, 256), 0, 256); w2 = rnd(256, 2304); o2 = ops.new_map(8, 40, 40, 256, dev); b2 = torch.zeros(256, device=dev)
cases.append(("conv3x3 s1 8x40x40 256->256", lambda: ops.conv3x3(x2, w2, o2, 1, b2, None, ops.ACT_SILU)))
x3 = ops.Map(rnd(8, 160, 160, 64), 0, 64); w3 = rnd(64, 64); o3 = ops.new_map(8, 160, 160, 64, dev)
cases.append(("conv1x1 8x160x160 64->64", lambda: ops.conv1x1(x3, w3, o3, b, None, ops.ACT_SILU)))
x4 = ops.Map(rnd(8, 160, 160, 32), 0, 32); w4 = rnd(32, 288); o4 = ops.new_map(8, 160, 160, 32, dev); b4 = torch.zeros(32, device=dev)
cases.append(("conv3x3 s1 8x160x160 32->32", lambda: ops.conv3x3(x4, w4, o4, 1, b4, None, ops.ACT_SILU)))
a5 = rnd(16384, 3072); w5 = rnd(1024, 3072); 
cases.append(("gemm 16384x1024x3072 f32 out", lambda: ops.linear(a5, w5, None, None, 0, torch.float32)))
for name, f in cases:
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us")
