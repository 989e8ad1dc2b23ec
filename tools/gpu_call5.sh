set -x
mkdir -p gpurun_out
export FMMT_DEBUG=1
timeout 300 python -m pytest tests/test_op_gemm.py -x -q -k "pair" 2>&1 | tail -5
timeout 300 python tests/gpu_gemm_pair_probe.py 2>&1 | tee gpurun_out/gemm_pair_probe.txt
