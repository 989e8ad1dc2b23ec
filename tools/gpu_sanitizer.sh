#!/bin/bash
# compute-sanitizer over the operator-level parity tests (one GPU, under gpurun): memcheck (out-of-bounds / misaligned accesses)
# and racecheck (shared-memory hazards between generic-proxy accesses; the TMA / tcgen05 async-proxy accesses are ordered
# by mbarriers and proxy fences, which racecheck does not model). Logs in gpurun_out/.
mkdir -p gpurun_out
SEL='gemm_shapes and (128-96-96 or 1568-576-192 or 245-768-3072 or 130-96-48 or 1-768-768) or gemm_epilogues or gemm_block_n or gemm_layernorm_epilogue and (3136 or 1000 or 260) or layernorm_plain or layernorm_window or window_attention or test_mha or fused_attention_half_block and (1-0 or 1-3) or ln_qkv_window_gather and 1-0 or swin_mlp or mlp_stream or span or filter_pack or frame_ingest'
T="tests/test_op_gemm.py tests/test_ops_gpu.py tests/test_attn_fused_gpu.py tests/test_ln_qkv_gpu.py tests/test_ingest_gpu.py"
for tool in memcheck racecheck; do
  timeout 540 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest $T -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool exit code $?" | tee -a gpurun_out/r02_sanitizer_$tool.log
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/r02_sanitizer_$tool.log | tail -8
done
