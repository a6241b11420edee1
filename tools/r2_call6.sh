#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_facade.py 2>&1 | grep -v Warn | tail -3
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c6_gpu_tests.log 2>&1; tail -15 gpurun_out/r2c6_gpu_tests.log
