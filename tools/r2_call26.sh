#!/bin/bash
# final single-GPU validation of the round: full GPU suite, smoke, the driver's bench command, reference arm, per-op table,
# ncu launch list + DRAM traffic of one parse step
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2c26_gpu_tests.log 2>&1; tail -4 gpurun_out/r2c26_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err; grep "leg\|verify\|caption stages\|device memory\|host threads" gpurun_out/r2c26_bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2c26_bench_ref.json 2> gpurun_out/r2c26_bench_ref.err; head -c 600 gpurun_out/r2c26_bench_ref.json; echo
timeout 400 python tools/time_ops.py all 416 > gpurun_out/r2c26_ops.txt 2> gpurun_out/r2c26_ops.err; grep "^==" gpurun_out/r2c26_ops.txt
B2P_NO_GRAPH=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c26_step_traffic.csv python tools/profile_step.py 2> gpurun_out/r2c26_prof.err; tail -2 gpurun_out/r2c26_prof.err
python tools/stage_traffic.py gpurun_out/r2c26_step_traffic.csv > gpurun_out/r2c26_stage_traffic.json; cat gpurun_out/r2c26_stage_traffic.json
