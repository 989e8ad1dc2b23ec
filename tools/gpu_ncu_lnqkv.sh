#!/bin/bash
# ncu full capture of the LN + qkv kernels (C=192: launches 0-1, C=384: launches 8-9) and the stage 2/3 window attention
export PROFILE_STEPS=1
P="python tools/profile_step.py"
F="--set full --clock-control none --import-source on"
ncu $F -k regex:ln_qkv_stream -c 2 -f -o gpurun_out/r02_ncu_lnqkv192 $P > gpurun_out/ncu_a.log 2>&1
ncu $F -k regex:ln_qkv_stream -s 8 -c 2 -f -o gpurun_out/r02_ncu_lnqkv384 $P > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
ls -la gpurun_out/*.ncu-rep
