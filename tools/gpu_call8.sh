mkdir -p gpurun_out
for c in 1 4 16 74; do echo "== COPIES=$c"; COPIES=$c timeout 200 python tests/gpu_mlp_stream_probe.py 2>&1 | grep -E "fused|timeout 0x[1-9a-f]"; done
