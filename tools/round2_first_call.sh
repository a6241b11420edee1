#!/bin/bash
# One GPU call that measures everything written at the end of round 1 without hardware access (see DESIGN.md §6,
# profiles/r1_gemm_notes.md "Next").  Usage:  gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Every step runs under its own timeout and writes to gpurun_out/; a failing step does not stop the others (the bounded
# mbarrier / split-K waits trap after ~2 s instead of hanging).
mkdir -p gpurun_out
# 1. grouped captioning: numerics, then throughput at group 2 with 2 and 1 lanes
B2P_TEST_UNVALIDATED=1 timeout 400 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -k "grouped or 3240" > gpurun_out/r2_group_test.log 2>&1
tail -3 gpurun_out/r2_group_test.log
for L in 2 1; do
  timeout 150 python bench.py --no-cpu-baseline --caption-group 2 --caption-lanes $L > gpurun_out/r2_bench_group2_l$L.json 2> gpurun_out/r2_bench_group2_l$L.err
  grep "leg:\|verify" gpurun_out/r2_bench_group2_l$L.err
done
# 2. CTA-pair GEMM: numerics (the debug log must show "b2p_gemm PAIR"), then the bench with it
B2P_CTA2=1 B2P_DEBUG=1 timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k cta_pair > gpurun_out/r2_cta2_test.log 2>&1
grep -c "b2p_gemm PAIR" gpurun_out/r2_cta2_test.log; tail -3 gpurun_out/r2_cta2_test.log
if grep -q " passed" gpurun_out/r2_cta2_test.log && ! grep -q "failed\|error" gpurun_out/r2_cta2_test.log; then
  B2P_CTA2=1 timeout 150 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_cta2.json 2> gpurun_out/r2_bench_cta2.err
  grep "leg:\|verify\|caption stages" gpurun_out/r2_bench_cta2.err
  B2P_CTA2=1 timeout 100 python tools/prof_gemm.py > gpurun_out/r2_prof_gemm_cta2.log 2>&1; tail -8 gpurun_out/r2_prof_gemm_cta2.log
fi
# 2b. smem-tiled dwconv+LN and warp-per-group channel attention: bit-identical with the validated kernels?  then the bench with both
B2P_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "dwconv_ln_tiled or channel_attn_small" > gpurun_out/r2_dwtile_test.log 2>&1; tail -2 gpurun_out/r2_dwtile_test.log
B2P_DWCONV_TILE=1 B2P_CHATTN_SMALL=1 timeout 150 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_dwtile.json 2> gpurun_out/r2_bench_dwtile.err; grep "leg:\|verify\|caption stages" gpurun_out/r2_bench_dwtile.err
# 3. where the microseconds of a GEMM launch go (instrumented kernel instantiation)
B2P_TRACE=1 B2P_NO_GRAPH=1 timeout 240 python tools/trace_gemm.py step > gpurun_out/r2_trace_step.txt 2>&1
tail -45 gpurun_out/r2_trace_step.txt
