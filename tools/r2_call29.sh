#!/bin/bash
# schedule / tiling-model variants on one box (defaults: lanes 2, group 2, sparse-grid penalty 0.04 us per CTA)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 env "$@" python bench.py --no-cpu-baseline $EXTRA > gpurun_out/r2c29_$tag.json 2> gpurun_out/r2c29_$tag.err; echo "== $tag: $(grep 'resident leg\|e2e leg:' gpurun_out/r2c29_$tag.err | tr '\n' ' ') $(grep 'device memory' gpurun_out/r2c29_$tag.err | tail -1)"; }
EXTRA="" run default A=1
EXTRA="--caption-group 3 --caption-lanes 2" run g3l2 A=1
EXTRA="--caption-group 4 --caption-lanes 1" run g4l1 A=1
EXTRA="" run dense B2P_DENSE_GRIDS=1
EXTRA="" run pen002 B2P_CTA_PENALTY=0.02
EXTRA="" run pen008 B2P_CTA_PENALTY=0.08
