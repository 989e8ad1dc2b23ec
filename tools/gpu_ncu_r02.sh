#!/bin/bash
# ncu captures of the round-2 build (one GPU; run under gpurun). Outputs in gpurun_out/.
export PROFILE_STEPS=1
P="python tools/profile_step.py"
F="--set full --clock-control none --import-source on"
L="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none"
ncu $F -k regex:swin_attn96 -c 2 -f -o gpurun_out/r02_ncu_attn96 $P > gpurun_out/ncu1.log 2>&1
ncu $F -k regex:swin_mlp_stream -s 8 -c 1 -f -o gpurun_out/r02_ncu_mlp384 $P > gpurun_out/ncu2.log 2>&1
ncu $F -k regex:swin_mlp96 -c 1 -f -o gpurun_out/r02_ncu_mlp96 $P > gpurun_out/ncu3.log 2>&1
ncu $F -k regex:window_attention_kernel -c 3 -f -o gpurun_out/r02_ncu_winattn $P > gpurun_out/ncu4.log 2>&1
ncu $L -k regex:gemm_bf16_tcgen05_tma -c 220 -f -o gpurun_out/r02_ncu_gemm $P > gpurun_out/ncu5.log 2>&1
ncu $L -k regex:"layernorm_vec|mha_flash|swin_mlp_stream" -c 200 -f -o gpurun_out/r02_ncu_ln_mha $P > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu5.log
ls -la gpurun_out/*.ncu-rep
