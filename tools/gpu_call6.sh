set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_op_gemm.py -x -q 2>&1 | tail -3
timeout 200 python tests/gpu_gemm_single_probe.py 2>&1 | tee gpurun_out/gemm_single_2producers.txt
timeout 300 python tests/gpu_gemm_pair_probe.py 2>&1 | tee gpurun_out/gemm_pair_probe.txt
