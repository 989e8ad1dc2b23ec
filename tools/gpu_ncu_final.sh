#!/bin/bash
# ncu evidence of the FINAL build (one GPU, under gpurun). Outputs in gpurun_out/; summarised into profiles/ with
# tools/ncu_summary.py. A number printed under ncu is never a bench value.
mkdir -p gpurun_out
P="python tools/profile_step.py"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
# 1. whole-step metric pass: two steps (PROFILE_STEPS=2), the second one is the summarised one
PROFILE_STEPS=2 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_final_launches.csv $P > gpurun_out/r02_final_launches.log 2>&1
tail -3 gpurun_out/r02_final_launches.log
# 2. --set full of one launch of each top kernel
export PROFILE_STEPS=1
F="--set full --clock-control none --import-source on"
ncu $F -k regex:swin_mlp_stream -s 8 -c 1 -f -o gpurun_out/r02_final_mlp384 $P > gpurun_out/ncu_f1.log 2>&1
ncu $F -k regex:swin_mlp_stream -s 1 -c 1 -f -o gpurun_out/r02_final_mlp192 $P > gpurun_out/ncu_f2.log 2>&1
ncu $F -k regex:swin_attn96 -s 1 -c 1 -f -o gpurun_out/r02_final_attn96 $P > gpurun_out/ncu_f3.log 2>&1
ncu $F -k regex:swin_mlp96 -s 1 -c 1 -f -o gpurun_out/r02_final_mlp96 $P > gpurun_out/ncu_f4.log 2>&1
ncu $F -k regex:ln_qkv_stream -s 5 -c 1 -f -o gpurun_out/r02_final_lnqkv384 $P > gpurun_out/ncu_f5.log 2>&1
ncu $F -k regex:ln_qkv_stream -s 1 -c 1 -f -o gpurun_out/r02_final_lnqkv192 $P > gpurun_out/ncu_f6.log 2>&1
ncu $F -k regex:window_attention_kernel -c 12 -f -o gpurun_out/r02_final_winattn $P > gpurun_out/ncu_f7.log 2>&1
ncu $F -k regex:mha_flash -c 2 -f -o gpurun_out/r02_final_mha $P > gpurun_out/ncu_f8.log 2>&1
ls -la gpurun_out/r02_final_*
