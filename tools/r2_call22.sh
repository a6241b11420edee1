#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "mha_short or elementwise or gemm" > gpurun_out/r2c22_ops.log 2>&1; tail -3 gpurun_out/r2c22_ops.log
timeout 600 python -m pytest tests/test_florence_gpu.py tests/test_yolo_gpu.py -m gpu -q -x > gpurun_out/r2c22_models.log 2>&1; tail -2 gpurun_out/r2c22_models.log
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c22_ops.txt 2> gpurun_out/r2c22_ops.err; grep "^==" gpurun_out/r2c22_ops.txt | grep "sum\|gemm\|conv3x3\|mha\|adown"; tail -3 gpurun_out/r2c22_ops.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c22_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 0 -c 10 -f -o gpurun_out/r2c22_gemm python tools/prof_gemm2.py > gpurun_out/r2c22_labels.txt 2> gpurun_out/r2c22_ncu.err; tail -2 gpurun_out/r2c22_ncu.err
