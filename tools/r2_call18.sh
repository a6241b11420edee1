#!/bin/bash
# florence_simt.cu with bulk copies + detector pooling kernels (32-bit index math, 4x4 patch max-pool): tests, per-op cost, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "v3 or mha_short or elementwise" > gpurun_out/r2c18_ops.log 2>&1; tail -8 gpurun_out/r2c18_ops.log
timeout 900 python -m pytest tests/test_florence_gpu.py tests/test_yolo_gpu.py -m gpu -q -x > gpurun_out/r2c18_models.log 2>&1; tail -5 gpurun_out/r2c18_models.log
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c18_ops.txt 2> gpurun_out/r2c18_ops.err; grep "^==" gpurun_out/r2c18_ops.txt; tail -3 gpurun_out/r2c18_ops.err
timeout 300 python bench.py --no-cpu-baseline --caption-lanes 3 > gpurun_out/r2c18_bench_l3.json 2> gpurun_out/r2c18_bench_l3.err; echo LANES3; grep "leg\|verify\|caption stages" gpurun_out/r2c18_bench_l3.err
