set -x
mkdir -p gpurun_out
timeout 200 python tests/gpu_mlp_stream_probe.py 2>&1 | tee gpurun_out/mlp_stream_probe.txt
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "stream" 2>&1 | tail -8
