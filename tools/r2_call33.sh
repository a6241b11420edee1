#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the same unit tests
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "v3 or mha_short or gemm_ln or multi_window" > gpurun_out/r2c33_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/r2c33_racecheck.log | head -20
