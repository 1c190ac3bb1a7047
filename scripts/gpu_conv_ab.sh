#!/bin/bash
# conv kernel A/B: parity tests, then the inference bench (conv table) with CTA pairs on and off
tag=${1:-ab}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_e2e.py tests/test_gpu_fp32.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15
for pairs in 1 0; do
  TEDM_CTA_PAIRS=$pairs TEDM_BENCH_CONV_TABLE=gpurun_out/${tag}_conv_table_pairs${pairs}.txt timeout 400 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline --no-fp32 > gpurun_out/${tag}_bench_pairs${pairs}.json 2> gpurun_out/${tag}_pairs${pairs}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_pairs${pairs}.json").read().strip().splitlines()[-1])
    print("pairs=${pairs}", round(d["value"], 1), round(d["ms_per_step"], 3), "conv", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), round(d["roofline"]["conv_ms_per_step"], 3), d["clocks"]["reasons"])
except Exception as e:
    print("pairs=${pairs} failed", e); print(open("gpurun_out/${tag}_pairs${pairs}.err").read()[-1200:])
PY
done
paste -d'|' <(cut -d'|' -f1,2 gpurun_out/${tag}_conv_table_pairs1.txt | head -16) <(cut -d'|' -f2 gpurun_out/${tag}_conv_table_pairs0.txt | head -16)
