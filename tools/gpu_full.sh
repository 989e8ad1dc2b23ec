# full GPU validation: parity tests + bench (+ per-kernel event profile)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernel_events.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
