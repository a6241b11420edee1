#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r2c12_ops.log 2>&1; tail -3 gpurun_out/r2c12_ops.log
timeout 200 python tools/time_yolo.py 1 8 2>&1 | grep -v Warn | tee gpurun_out/r2c12_time_yolo.log
timeout 900 python -m pytest tests/test_yolo_gpu.py tests/test_florence_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r2c12_models.log 2>&1; tail -3 gpurun_out/r2c12_models.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c12_bench.err
B2P_TRACE=1 B2P_NO_GRAPH=1 timeout 300 python tools/trace_gemm.py step > gpurun_out/r2c12_trace_step.txt 2>&1; tail -42 gpurun_out/r2c12_trace_step.txt | cut -c1-150
