#!/bin/bash
# attribution on the final tree, one box, back to back: each round-2 change switched off in turn (bench.py defaults otherwise)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 env "$@" python bench.py --no-cpu-baseline $EXTRA > gpurun_out/r2c31_$tag.json 2> gpurun_out/r2c31_$tag.err; echo "== $tag: $(grep 'resident leg\|e2e leg:\|caption stages' gpurun_out/r2c31_$tag.err | sed 's/\[bench *[0-9.]*s\] //' | tr '\n' ';')"; }
EXTRA="" run final A=1
EXTRA="" run no_simt_v3 B2P_NO_SIMT_V3=1
EXTRA="" run no_forced_skip B2P_NO_FORCED_SKIP=1
EXTRA="--caption-lanes 2 --caption-group 1" run lanes2_group1 A=1
EXTRA="" run pdl_simt B2P_PDL_SIMT=1
EXTRA="" run final_again A=1
