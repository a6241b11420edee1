#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r2c4_ops.log 2>&1; tail -3 gpurun_out/r2c4_ops.log
timeout 600 python -m pytest tests/test_yolo_gpu.py tests/test_florence_gpu.py -m gpu -q -x -s > gpurun_out/r2c4_models.log 2>&1; tail -3 gpurun_out/r2c4_models.log; grep "x3 head\|boxes, max\|tie-class" gpurun_out/r2c4_models.log
timeout 900 python -m pytest tests/test_boundary_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s > gpurun_out/r2c4_pipe.log 2>&1; tail -5 gpurun_out/r2c4_pipe.log; grep "integer-boundary" gpurun_out/r2c4_pipe.log
timeout 100 python tools/prof_gemm.py > gpurun_out/r2c4_prof_bres.log 2>&1; head -5 gpurun_out/r2c4_prof_bres.log
B2P_NO_BRES=1 timeout 100 python tools/prof_gemm.py > gpurun_out/r2c4_prof_nobres.log 2>&1; head -5 gpurun_out/r2c4_prof_nobres.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; grep "leg\|verify\|caption stages" gpurun_out/r2c4_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c4_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"fwd_ms",d["roofline"]["forward_ms"],"frac",d["roofline"]["frac"],"p50",d["p50_latency_ms_batch1"])
PY
B2P_NO_BRES=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c4_bench_nobres.json 2> gpurun_out/r2c4_bench_nobres.err; grep "leg\|caption stages" gpurun_out/r2c4_bench_nobres.err
