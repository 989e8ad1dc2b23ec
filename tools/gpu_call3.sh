set -x
mkdir -p gpurun_out
timeout 60 ./tools/bench_gelu 2>&1 | tee gpurun_out/gelu_bench.txt
timeout 300 python tests/gpu_mlp_probe.py 2>&1 | tee gpurun_out/mlp_probe.txt
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_swin_gpu.py -x -q 2>&1 | tail -15
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1800 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"swin_mlp96|window_attention" -c 6 -o gpurun_out/r01_s4_mlp_wa python tests/gpu_profile_swin.py 64 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
ncu -i gpurun_out/r01_s4_mlp_wa.ncu-rep --page raw --csv > gpurun_out/r01_s4_mlp_wa_raw.csv 2>/dev/null
ls -la gpurun_out
