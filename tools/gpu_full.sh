# full GPU validation: parity tests, bench (+ per-kernel event profile), ncu launch list of one bench step, ncu full of the top kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; head -c 330 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_bf16_tcgen05_tma|swin_mlp|window_attention" -c 16 -o gpurun_out/r01_final_swin python tests/gpu_profile_swin.py 160 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/r01_final_swin.ncu-rep --page raw --csv > gpurun_out/r01_final_swin_raw.csv 2>/dev/null
ls -la gpurun_out | head -20
