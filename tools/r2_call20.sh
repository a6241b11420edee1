#!/bin/bash
# dwconv_ln v3 (zero-pixel select, rolled loops), rolled mha_short, batched cbfuse: tests, per-op cost, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "v3 or mha_short or elementwise" > gpurun_out/r2c20_ops.log 2>&1; tail -4 gpurun_out/r2c20_ops.log
timeout 900 python -m pytest tests/test_florence_gpu.py tests/test_yolo_gpu.py -m gpu -q -x > gpurun_out/r2c20_models.log 2>&1; tail -3 gpurun_out/r2c20_models.log
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c20_ops.txt 2> gpurun_out/r2c20_ops.err; grep "^==" gpurun_out/r2c20_ops.txt; tail -3 gpurun_out/r2c20_ops.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c20_bench.err
