#!/bin/bash
# re-entry baseline: bench line of the committed tree + live per-op cost table (caption plan K=416, detector batch 8)
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c16_bench.err
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c16_ops.txt 2> gpurun_out/r2c16_ops.err; grep "^==" gpurun_out/r2c16_ops.txt; tail -3 gpurun_out/r2c16_ops.err
timeout 300 python bench.py --no-cpu-baseline --caption-lanes 3 > gpurun_out/r2c16_bench_l3.json 2> gpurun_out/r2c16_bench_l3.err; echo LANES3; grep "leg\|verify\|caption stages" gpurun_out/r2c16_bench_l3.err | head; tail -2 gpurun_out/r2c16_bench_l3.err
