#!/bin/bash
# last single-GPU sanity of the final tree (im2col_u8 reverted to the per-pixel kernel): image-op tests, smoke, the driver's bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_yolo_gpu.py -m gpu -q -k "im2col or stem or yolo or predict" > gpurun_out/r2c28_ops.log 2>&1; tail -2 gpurun_out/r2c28_ops.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err; grep "leg\|verify\|caption stages\|device memory" gpurun_out/r2c28_bench.err
