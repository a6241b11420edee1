#!/bin/bash
# b2p_gemm_ln (park-only split-K GEMM + reduce/LayerNorm), batched greedy pick, mha split reverted: tests, per-op cost, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm_ln or mha_short" > gpurun_out/r2c23_ops.log 2>&1; tail -3 gpurun_out/r2c23_ops.log
timeout 900 python -m pytest tests/test_florence_gpu.py tests/test_pipeline_gpu.py tests/test_boundary_gpu.py -m gpu -q -x > gpurun_out/r2c23_models.log 2>&1; tail -3 gpurun_out/r2c23_models.log
timeout 400 python tools/time_ops.py florence 416 > gpurun_out/r2c23_ops.txt 2> gpurun_out/r2c23_ops.err; grep "^== dec\|^== enc: \|== enc mha" gpurun_out/r2c23_ops.txt; tail -3 gpurun_out/r2c23_ops.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c23_bench.err
B2P_NO_GEMM_LN=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c23_bench_noln.json 2> gpurun_out/r2c23_bench_noln.err; echo NO_GEMM_LN; grep "resident\|caption stages" gpurun_out/r2c23_bench_noln.err
