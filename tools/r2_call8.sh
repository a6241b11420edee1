#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_determinism.py 2>&1 | grep -v Warn | tail -14
timeout 600 python -m pytest tests/test_boundary_gpu.py -m gpu -q -s -k "real or batching" > gpurun_out/r2c8_bound.log 2>&1; tail -4 gpurun_out/r2c8_bound.log; grep "order events\|integer-boundary" gpurun_out/r2c8_bound.log
PROF_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 4 -o gpurun_out/r2c8_prof python tools/prof_gemm.py > gpurun_out/r2c8_prof.log 2>&1; tail -3 gpurun_out/r2c8_prof.log; ls -la gpurun_out/r2c8_prof.ncu-rep
