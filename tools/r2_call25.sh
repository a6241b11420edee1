#!/bin/bash
# plan buffer pool (DaViT stages + BART encoder re-use their intermediates), bias prefetch + multi-window QB=2 reverted
mkdir -p gpurun_out
export B2P_BUFFER_POOL=1
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "window or v3" > gpurun_out/r2c25_ops.log 2>&1; tail -2 gpurun_out/r2c25_ops.log
timeout 1200 python -m pytest tests/test_florence_gpu.py tests/test_pipeline_gpu.py tests/test_boundary_gpu.py -m gpu -q -x > gpurun_out/r2c25_models.log 2>&1; tail -3 gpurun_out/r2c25_models.log
timeout 100 python tools/time_yolo.py 8 2>&1 | grep "graph=True"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err; grep "resident\|e2e leg:\|caption stages\|device memory\|verify" gpurun_out/r2c25_bench.err
timeout 300 python bench.py --no-cpu-baseline --caption-group 2 --caption-lanes 2 > gpurun_out/r2c25_bench_g2l2.json 2> gpurun_out/r2c25_bench_g2l2.err; echo G2L2; grep "resident\|e2e leg:\|verify\|device memory" gpurun_out/r2c25_bench_g2l2.err; tail -1 gpurun_out/r2c25_bench_g2l2.err
timeout 300 python bench.py --no-cpu-baseline --caption-group 2 --caption-lanes 3 > gpurun_out/r2c25_bench_g2l3.json 2> gpurun_out/r2c25_bench_g2l3.err; echo G2L3; grep "resident\|e2e leg:\|verify\|device memory" gpurun_out/r2c25_bench_g2l3.err; tail -1 gpurun_out/r2c25_bench_g2l3.err
timeout 300 python bench.py --no-cpu-baseline --caption-lanes 4 > gpurun_out/r2c25_bench_l4.json 2> gpurun_out/r2c25_bench_l4.err; echo L4; grep "resident\|e2e leg:\|verify\|device memory" gpurun_out/r2c25_bench_l4.err; tail -1 gpurun_out/r2c25_bench_l4.err
