#!/bin/bash
# N = 2 (torchrun, one rank per GPU over NCCL): the driver's scaling command
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c27_bench_n2.json 2> gpurun_out/r2c27_bench_n2.err; grep "leg\|verify\|caption stages" gpurun_out/r2c27_bench_n2.err | head; head -c 400 gpurun_out/r2c27_bench_n2.json; echo
