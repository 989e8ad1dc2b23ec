set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_op_gemm.py -x -q 2>&1 | tail -3
for k in 1 0 2; do FMMT_KBS=$k timeout 200 python tests/gpu_gemm_single_probe.py 2>&1 | tee gpurun_out/gemm_single_kbs$k.txt; done
