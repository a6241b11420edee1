#!/bin/bash
# compute-sanitizer memcheck over the unit tests of the kernels written in the second half of round 2
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "v3 or mha_short or gemm_ln or elementwise or multi_window" > gpurun_out/r2c32_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r2c32_memcheck.log | head -20
