"""Representative YOLOv9-E GEMM/conv shapes at batch 8 for `ncu --set full` (profiles/): 
  ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 8 -o gpurun_out/prof_gemm python tools/prof_gemm.py   (PROF_ONCE=1)"""
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
# fp16x3 operand mode (caption path): [hi | lo] operands, three MMAs per k-step
def hilo(t):
    hi = t.half(); return torch.cat([hi, (t - hi.float()).half()], 1).contiguous().to(dev)
a6 = hilo(torch.randn(6656, 512, generator=g)); w6 = hilo(torch.randn(2048, 512, generator=g) * 0.05); o6 = torch.empty(6656, 2 * 2048, dtype=torch.float16, device=dev); b6 = torch.zeros(2048, device=dev)
cases.append(("x3 gemm 6656x2048x512 gelu split-out (DaViT stage-2 fc1)", lambda: ops.gemm(a6, 1024, w6, 6656, 2048, 512, o6, 4096, b6, None, 0, ops.ACT_GELU, split=True, x3=True)))
a7 = hilo(torch.randn(416, 768, generator=g)); w7 = hilo(torch.randn(768, 768, generator=g) * 0.05); o7 = torch.empty(416, 768, dtype=torch.float32, device=dev); r7 = torch.zeros(416, 768, device=dev); b7 = torch.zeros(768, device=dev)
cases.append(("x3 gemm 416x768x768 f32+res (decoder out-proj, split-K)", lambda: ops.gemm(a7, 1536, w7, 416, 768, 768, o7, 768, b7, r7, 768, 0, out_f32=True, x3=True)))
w8 = hilo(torch.randn(51290, 768, generator=g) * 0.05); o8 = torch.empty(416, 51290, dtype=torch.float32, device=dev)
cases.append(("x3 gemm 416x51290x768 f32 (LM head)", lambda: ops.gemm(a7, 1536, w8, 416, 51290, 768, o8, 51290, None, None, 0, 0, out_f32=True, x3=True)))
import os
if os.environ.get("PROF_ONCE"):   # under `ncu --set full`: every case exactly once (keeps the report small)
    for name, f in cases:
        f()
        torch.cuda.synchronize()
        print(name)
    sys.exit(0)
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
