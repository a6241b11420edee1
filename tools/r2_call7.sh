#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k overlap > gpurun_out/r2c7_ovl.log 2>&1; tail -4 gpurun_out/r2c7_ovl.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r2c7_pipe.log 2>&1; tail -4 gpurun_out/r2c7_pipe.log
timeout 900 python -m pytest tests/test_boundary_gpu.py tests/test_yolo_gpu.py -m gpu -q -s > gpurun_out/r2c7_bound.log 2>&1; tail -8 gpurun_out/r2c7_bound.log; grep "tie-class\|order events\|integer-boundary\|boxes, max" gpurun_out/r2c7_bound.log
B2P_NO_BRES=1 timeout 200 python tools/time_yolo.py 8 2>&1 | grep -v Warn | tee gpurun_out/r2c7_time_yolo_nobres.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c7_bench.err
B2P_HOST_GLUE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c7_bench_hostglue.json 2> gpurun_out/r2c7_bench_hostglue.err; grep "leg\|verify" gpurun_out/r2c7_bench_hostglue.err
B2P_TRACE=1 B2P_NO_GRAPH=1 timeout 300 python tools/trace_gemm.py step > gpurun_out/r2c7_trace_step.txt 2>&1; tail -45 gpurun_out/r2c7_trace_step.txt
