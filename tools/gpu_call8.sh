mkdir -p gpurun_out
COPIES=1 timeout 200 python tests/gpu_mlp_stream_probe.py 2>&1 | grep -E "fused|timeout 0x[1-9a-f]" | tee gpurun_out/mlp_stream_probe.txt
timeout 200 python -m pytest tests/test_ops_gpu.py -x -q -k "stream" 2>&1 | tail -3
