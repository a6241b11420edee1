#!/bin/bash
# bias prefetch before the TMEM wait; caption grouping re-measured with the round-2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm or conv or window or v3" > gpurun_out/r2c24_ops.log 2>&1; tail -3 gpurun_out/r2c24_ops.log
timeout 900 python -m pytest tests/test_florence_gpu.py tests/test_yolo_gpu.py -m gpu -q -x > gpurun_out/r2c24_models.log 2>&1; tail -2 gpurun_out/r2c24_models.log
timeout 300 python tools/time_ops.py florence 416 2>/dev/null | grep "window_attn\|^== enc" | sort | uniq -c | sort -k5 -n -r | head -12
timeout 100 python tools/time_yolo.py 8 2>&1 | grep "graph=True"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err; grep "resident\|e2e leg:\|caption stages" gpurun_out/r2c24_bench.err
timeout 300 python bench.py --no-cpu-baseline --caption-group 2 --caption-lanes 2 > gpurun_out/r2c24_bench_g2l2.json 2> gpurun_out/r2c24_bench_g2l2.err; echo G2L2; grep "resident\|e2e leg:\|verify" gpurun_out/r2c24_bench_g2l2.err; tail -2 gpurun_out/r2c24_bench_g2l2.err
timeout 300 python bench.py --no-cpu-baseline --caption-group 2 --caption-lanes 3 > gpurun_out/r2c24_bench_g2l3.json 2> gpurun_out/r2c24_bench_g2l3.err; echo G2L3; grep "resident\|e2e leg:\|verify" gpurun_out/r2c24_bench_g2l3.err; tail -2 gpurun_out/r2c24_bench_g2l3.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
