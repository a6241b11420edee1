#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_overlap.py 2>&1 | grep -v Warn | tail -6
timeout 900 python tools/sweep_decode_gemm.py 2>&1 | grep -v Warn | tee gpurun_out/r2c9_sweep.log | tail -90
