set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json
timeout 120 python tests/gpu_gemm_perf.py > gpurun_out/gemm_perf.txt 2>&1; cat gpurun_out/gemm_perf.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off -c 75 -o gpurun_out/r01_swin64_full python tests/gpu_profile_swin.py 64 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
ls -la gpurun_out
