#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r2c10_ops.log 2>&1; tail -4 gpurun_out/r2c10_ops.log
timeout 200 python tools/time_yolo.py 1 8 2>&1 | grep -v Warn | tee gpurun_out/r2c10_time_yolo.log
timeout 900 python -m pytest tests/test_yolo_gpu.py tests/test_florence_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r2c10_models.log 2>&1; tail -4 gpurun_out/r2c10_models.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c10_bench.err
timeout 200 python tools/sweep_decode_gemm.py 2>&1 | grep -v Warn | grep "bn_max 256 auto\|---" | tee gpurun_out/r2c10_sweep.log
