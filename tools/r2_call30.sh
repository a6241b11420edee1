#!/bin/bash
# sparse-grid charge 0.02 us per CTA as the default: full GPU suite + the driver's bench on the final tree
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2c30_gpu_tests.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2c30_gpu_tests.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err; grep "leg\|verify\|caption stages\|device memory" gpurun_out/r2c30_bench.err
