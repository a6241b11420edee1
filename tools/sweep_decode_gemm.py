"""Tiling sweep of the decode-step GEMMs (M = 416 rows, fp16x3): for each shape, every (N tile cap, split-K on/off) timed as
it runs inside the decode step -- a CUDA graph of 24 back-to-back launches on one stream (PDL chain) -- with CUDA events.
  python tools/sweep_decode_gemm.py [M]"""
import sys
sys.path.insert(0, ".")
import torch
from omniparser_b200 import ops

dev = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 416
g = torch.Generator().manual_seed(0)


def hilo(t):
    hi = t.half()
    return torch.cat([hi, (t - hi.float()).half()], 1).contiguous().to(dev)


SHAPES = [("o/co/cq 768x768", 768, 768, dict(out_f32=True, res=True)), ("qkv 2304x768", 2304, 768, dict(out_f32=True)),
          ("fc1 3072x768 gelu split", 3072, 768, dict(act=ops.ACT_GELU, split=True)), ("fc2 768x3072", 768, 3072, dict(out_f32=True, res=True)),
          ("lm head 51290x768", 51290, 768, dict(out_f32=True))]
for name, N, K, cfg in SHAPES:
    a = hilo(torch.randn(M, K, generator=g))
    w = hilo(torch.randn(N, K, generator=g) * 0.05)
    bias = torch.zeros(N, device=dev) if N < 50000 else None
    if cfg.get("split"):
        out = torch.empty(M, 2 * N, dtype=torch.float16, device=dev); ldc = 2 * N
    else:
        out = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.float32, device=dev); ldc = out.stride(0)
    res = torch.zeros(M, N, device=dev) if cfg.get("res") else None
    print(f"--- {name}  (M={M})")
    for bn in (16, 32, 48, 64, 96, 128, 192, 256):
        for nosplit in (False, True):
            def f():
                ops.gemm(a, a.stride(0), w, M, N, K, out, ldc, bias, res, N if res is not None else 0, cfg.get("act", 0),
                         out_f32=cfg.get("out_f32", False), split=cfg.get("split", False), x3=True, bn_max=bn, no_splitk=nosplit)
            f(); torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(24):
                    f()
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"  bn_max {bn:3d} {'no-split' if nosplit else 'auto    '}: {e0.elapsed_time(e1) / 120 * 1e3:6.1f} us/launch", flush=True)
