# full GPU validation: parity tests, bench (+ per-kernel event profile), ncu launch list, ncu full capture of the top kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1700 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; tail -c 600 gpurun_out/bench_ref.json
