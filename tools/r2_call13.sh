#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r2c13_ops.log 2>&1; tail -3 gpurun_out/r2c13_ops.log
timeout 900 python -m pytest tests/test_yolo_gpu.py tests/test_florence_gpu.py -m gpu -q -x > gpurun_out/r2c13_models.log 2>&1; tail -3 gpurun_out/r2c13_models.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c13_bench.err
B2P_FEW_CTAS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c13_bench_few.json 2> gpurun_out/r2c13_bench_few.err; echo FEW_CTAS; grep "leg\|verify\|caption stages" gpurun_out/r2c13_bench_few.err
timeout 200 python tools/sweep_decode_gemm.py 2>&1 | grep -v Warn | grep "bn_max 256 auto\|---" | tee gpurun_out/r2c13_sweep.log
B2P_FEW_CTAS=1 B2P_DEBUG=1 timeout 200 python tools/sweep_decode_gemm.py 2>&1 | grep -v Warn | grep "bn_max 256 auto\|---\|b2p_gemm" | sort -u | tee gpurun_out/r2c13_sweep_few.log | tail -20
