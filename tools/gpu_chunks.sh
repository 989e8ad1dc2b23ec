mkdir -p gpurun_out
for cfg in "640 1280" "1280 1280" "320 1280" "640 640"; do set -- $cfg; echo -n "chunk $1 late $2: "; timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --swin-chunk $1 --swin-chunk-late $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; done | tee gpurun_out/chunk_sweep3.txt
