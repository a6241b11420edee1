"""Four epilogue-heavy shapes, once each, for `ncu --set full --import-source on -k regex:gemm_tcgen05` (round 2, after act16)."""
import sys
sys.path.insert(0, ".")
import torch
from omniparser_b200 import ops
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g) * 0.3).half().to(dev)
def hilo(t):
    hi = t.half(); return torch.cat([hi, (t - hi.float()).half()], 1).contiguous().to(dev)
cases = []
a1 = rnd(819200, 32); w1 = rnd(128, 32); o1 = torch.empty(819200, 128, dtype=torch.float16, device=dev); b1 = torch.zeros(128, device=dev)
cases.append(("stem gemm 819200x128x32 silu", lambda: ops.gemm(a1, 32, w1, 819200, 128, 32, o1, 128, b1, None, 0, ops.ACT_SILU)))
x4 = ops.Map(rnd(8, 160, 160, 32), 0, 32); w4 = rnd(32, 288); o4 = ops.new_map(8, 160, 160, 32, dev); b4 = torch.zeros(32, device=dev)
cases.append(("conv3x3 s1 8x160x160 32->32 silu", lambda: ops.conv3x3(x4, w4, o4, 1, b4, None, ops.ACT_SILU)))
x2 = ops.Map(rnd(8, 80, 80, 256), 0, 256); w2 = rnd(256, 2304); o2 = ops.new_map(8, 80, 80, 256, dev); b2 = torch.zeros(256, device=dev)
cases.append(("conv3x3 s1 8x80x80 256->256 silu", lambda: ops.conv3x3(x2, w2, o2, 1, b2, None, ops.ACT_SILU)))
a6 = hilo(torch.randn(6656, 512, generator=g)); w6 = hilo(torch.randn(2048, 512, generator=g) * 0.05); o6 = torch.empty(6656, 2 * 2048, dtype=torch.float16, device=dev); b6 = torch.zeros(2048, device=dev)
cases.append(("x3 gemm 6656x2048x512 gelu split-out", lambda: ops.gemm(a6, 1024, w6, 6656, 2048, 512, o6, 4096, b6, None, 0, ops.ACT_GELU, split=True, x3=True)))
a9 = hilo(torch.randn(6656, 2048, generator=g)); w9 = hilo(torch.randn(512, 2048, generator=g) * 0.05); o9 = torch.empty(6656, 512, dtype=torch.float32, device=dev); r9 = torch.zeros(6656, 512, device=dev); b9 = torch.zeros(512, device=dev)
cases.append(("x3 gemm 6656x512x2048 f32+res", lambda: ops.gemm(a9, 4096, w9, 6656, 512, 2048, o9, 512, b9, r9, 512, 0, out_f32=True, x3=True)))
for name, f in cases:
    f(); f()
    torch.cuda.synchronize()
    print(name)
