#!/bin/bash
# every single-GPU bench line of the round (BASELINE configs 2, 3, 4 per GPU, 5, the fp32-grade mode, the uint8 ingest)
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3"
$B --profile-out gpurun_out/r02_kernel_events_final.json > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
$B --ingest u8 --no-cpu-baseline > gpurun_out/r02_bench_n1_u8ingest.json 2>> gpurun_out/r02_bench_n1_final.err
$B --precision fp32 --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_n1_fp32mode.json 2>> gpurun_out/r02_bench_n1_final.err
$B --workload tav_bert_u32 --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_n1_bert_u32.json 2>> gpurun_out/r02_bench_n1_final.err
$B --workload tav_roberta_u32 --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_n1_roberta_u32.json 2>> gpurun_out/r02_bench_n1_final.err
$B --workload swin160 --no-cpu-baseline > gpurun_out/r02_bench_n1_swin160.json 2>> gpurun_out/r02_bench_n1_final.err
for f in gpurun_out/r02_bench_n1_final.json gpurun_out/r02_bench_n1_u8ingest.json gpurun_out/r02_bench_n1_fp32mode.json gpurun_out/r02_bench_n1_bert_u32.json gpurun_out/r02_bench_n1_roberta_u32.json gpurun_out/r02_bench_n1_swin160.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1].split('/')[-1], round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", d.get("e2e", {}).get("value"),
          "parity", d.get("parity_checked"), "frac", d.get("roofline", {}) and round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -5 gpurun_out/r02_bench_n1_final.err
