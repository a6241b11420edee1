#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/coop_test/coop_test 2>&1 | tail -5 | tee gpurun_out/r2c11_coop.log
for cfg in "2 1" "2 2" "3 1"; do set -- $cfg
  timeout 300 python bench.py --no-cpu-baseline --caption-group $1 --caption-lanes $2 > gpurun_out/r2c11_bench_g$1_l$2.json 2> gpurun_out/r2c11_bench_g$1_l$2.err; echo "group $1 lanes $2"; grep "leg:\|verify" gpurun_out/r2c11_bench_g$1_l$2.err
done
B2P_TRACE=1 B2P_NO_GRAPH=1 timeout 300 python tools/trace_gemm.py step > gpurun_out/r2c11_trace_step.txt 2>&1; tail -42 gpurun_out/r2c11_trace_step.txt
