# full GPU validation: parity tests, bench (+ per-kernel event profile)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json | head -c 400; tail -3 gpurun_out/bench_n1.err
FMMT_NO_MLP_STREAM=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_nostream.json 2>/dev/null; head -c 250 gpurun_out/bench_n1_nostream.json
