#!/bin/bash
# round-2 SIMT kernels (csrc/florence_simt.cu): unit tests vs the first versions + torch, caption parity, per-op cost, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "v3 or mha_short" > gpurun_out/r2c17_ops.log 2>&1; tail -15 gpurun_out/r2c17_ops.log
timeout 900 python -m pytest tests/test_florence_gpu.py -m gpu -q -x > gpurun_out/r2c17_florence.log 2>&1; tail -5 gpurun_out/r2c17_florence.log
timeout 400 python tools/time_ops.py florence 416 > gpurun_out/r2c17_ops.txt 2> gpurun_out/r2c17_ops.err; grep "^==" gpurun_out/r2c17_ops.txt; tail -3 gpurun_out/r2c17_ops.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c17_bench.err
timeout 300 python bench.py --no-cpu-baseline --caption-lanes 3 > gpurun_out/r2c17_bench_l3.json 2> gpurun_out/r2c17_bench_l3.err; echo LANES3; grep "leg\|verify\|caption stages" gpurun_out/r2c17_bench_l3.err
