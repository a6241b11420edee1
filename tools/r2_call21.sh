#!/bin/bash
# act16 (activation switch outside the epilogue element loop) + LM-head-free forced steps: full GPU suite, per-op cost, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c21_gpu_tests.log 2>&1; tail -6 gpurun_out/r2c21_gpu_tests.log
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c21_ops.txt 2> gpurun_out/r2c21_ops.err; grep "^==" gpurun_out/r2c21_ops.txt | grep "sum\|gemm\|conv3x3"; tail -3 gpurun_out/r2c21_ops.err
timeout 100 python tools/time_yolo.py 8 2>&1 | grep "graph=True"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c21_bench.err
