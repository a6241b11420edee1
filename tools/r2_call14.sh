#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c14_gpu_tests.log 2>&1; tail -6 gpurun_out/r2c14_gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c14_bench.err
B2P_FEW_CTAS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c14_bench_few.json 2> gpurun_out/r2c14_bench_few.err; echo FEW_CTAS; grep "leg\|verify\|caption stages" gpurun_out/r2c14_bench_few.err
B2P_SPLITK_LAST=1 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm or conv" > gpurun_out/r2c14_ops_sklast.log 2>&1; tail -2 gpurun_out/r2c14_ops_sklast.log
