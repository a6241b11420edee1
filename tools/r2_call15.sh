#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r2c15_ops.log 2>&1; tail -2 gpurun_out/r2c15_ops.log
timeout 600 python -m pytest tests/test_yolo_gpu.py tests/test_florence_gpu.py -m gpu -q -x > gpurun_out/r2c15_models.log 2>&1; tail -2 gpurun_out/r2c15_models.log
timeout 100 python tools/time_yolo.py 8 2>&1 | grep "graph=True"
B2P_NO_V256=1 timeout 100 python tools/time_yolo.py 8 2>&1 | grep "graph=True" | sed 's/^/NO_V256 /'
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; grep "resident\|caption stages" gpurun_out/r2c15_bench.err
B2P_NO_V256=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c15_bench_nov256.json 2> gpurun_out/r2c15_bench_nov256.err; echo NO_V256; grep "resident\|caption stages" gpurun_out/r2c15_bench_nov256.err
B2P_FEW_CTAS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c15_bench_few.json 2> gpurun_out/r2c15_bench_few.err; echo FEW_CTAS; grep "resident\|caption stages" gpurun_out/r2c15_bench_few.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c15_bench2.json 2> gpurun_out/r2c15_bench2.err; echo default again; grep "resident\|caption stages" gpurun_out/r2c15_bench2.err
