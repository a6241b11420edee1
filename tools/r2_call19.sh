#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv|window_attn|channel_attn|mha|adown|cbfuse|layernorm' -f -o gpurun_out/r2c19_simt python tools/prof_simt.py > gpurun_out/r2c19_labels.txt 2> gpurun_out/r2c19_ncu.err; tail -3 gpurun_out/r2c19_ncu.err; tail -20 gpurun_out/r2c19_labels.txt
ls -la gpurun_out/r2c19_simt.ncu-rep
