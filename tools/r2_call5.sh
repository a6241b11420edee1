#!/bin/bash
# round 2, call 5 (re-entry baseline): full GPU suite, bench, detector timing, per-launch list with shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c5_gpu_tests.log 2>&1; tail -5 gpurun_out/r2c5_gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c5_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c5_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"roofline",json.dumps(d["roofline"])[:600],"p50",d.get("p50_latency_ms_batch1"))
PY
timeout 200 python tools/time_yolo.py 1 8 2>&1 | grep -v Warn | tee gpurun_out/r2c5_time_yolo.log
B2P_NO_GRAPH=1 B2P_DEBUG=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c5_step_launches.csv python tools/profile_step.py 2> gpurun_out/r2c5_step_shapes.log; tail -2 gpurun_out/r2c5_step_shapes.log
