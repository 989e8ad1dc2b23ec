#!/bin/bash
# usage: tools/gpu_scale.sh N   (under `gpurun --gpus N`): headline workload + BASELINE configs[3] (U=32 per GPU) at N GPUs
N=$1
mkdir -p gpurun_out
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --warmup 3"
$R --steps 10 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
$R --steps 5 --workload tav_roberta_u32 > gpurun_out/r02_bench_n${N}_roberta_u32.json 2>> gpurun_out/r02_bench_n${N}.err
for f in gpurun_out/r02_bench_n${N}.json gpurun_out/r02_bench_n${N}_roberta_u32.json; do
python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1].split('/')[-1], round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2))
    print(json.dumps(d.get("scaling_detail")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/r02_bench_n${N}.err
