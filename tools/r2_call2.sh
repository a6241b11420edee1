#!/bin/bash
# round 2, call 2: parity-grade detector tests + scheduling experiments (split-K off, 3 caption lanes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_yolo_gpu.py -m gpu -q -x -s > gpurun_out/r2c2_yolo.log 2>&1; tail -40 gpurun_out/r2c2_yolo.log | grep -v Warning
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q > gpurun_out/r2c2_ops.log 2>&1; tail -3 gpurun_out/r2c2_ops.log
B2P_NO_SPLITK=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c2_bench_nosplit.json 2> gpurun_out/r2c2_bench_nosplit.err; grep "leg:\|verify\|caption stages" gpurun_out/r2c2_bench_nosplit.err
timeout 200 python bench.py --no-cpu-baseline --caption-lanes 3 > gpurun_out/r2c2_bench_l3.json 2> gpurun_out/r2c2_bench_l3.err; grep "leg:\|verify\|caption stages" gpurun_out/r2c2_bench_l3.err
B2P_NO_SPLITK=1 timeout 200 python bench.py --no-cpu-baseline --caption-lanes 3 > gpurun_out/r2c2_bench_nosplit_l3.json 2> gpurun_out/r2c2_bench_nosplit_l3.err; grep "leg:\|verify\|caption stages" gpurun_out/r2c2_bench_nosplit_l3.err
