#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c3_gpu_tests.log 2>&1; tail -5 gpurun_out/r2c3_gpu_tests.log
grep "x3 tap\|x3 head\|boxes, max\|tie-class" gpurun_out/r2c3_gpu_tests.log | tail -20
timeout 400 python bench.py --no-cpu-baseline --with-768 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; grep "leg\|verify\|caption stages\|768" gpurun_out/r2c3_bench.err; tail -c 600 gpurun_out/r2c3_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c3_ref.json 2> gpurun_out/r2c3_ref.err; tail -c 700 gpurun_out/r2c3_ref.json
B2P_NO_GRAPH=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_step_traffic.csv python tools/profile_step.py 2> gpurun_out/r2c3_prof.err; tail -2 gpurun_out/r2c3_prof.err
python tools/stage_traffic.py gpurun_out/r2_step_traffic.csv > gpurun_out/r2_stage_traffic.json; cat gpurun_out/r2_stage_traffic.json | head -40
