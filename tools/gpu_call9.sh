mkdir -p gpurun_out
timeout 250 python tests/gpu_gemm_stress.py 40 2>&1 | tail -8
