#!/bin/bash
# four-row weight-stationary conv tiles: parity tests, then the inference bench (conv table) with them on (TEDM_WS=1) and off (2)
tag=${1:-ws4}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_e2e.py tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15
for ws in 1 2; do
  TEDM_WS=$ws TEDM_BENCH_CONV_TABLE=gpurun_out/${tag}_conv_table_ws${ws}.txt timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/${tag}_bench_ws${ws}.json 2> gpurun_out/${tag}_ws${ws}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_ws${ws}.json").read().strip().splitlines()[-1])
    print("ws=${ws}", round(d["value"], 1), round(d["ms_per_step"], 3), "conv", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), round(d["roofline"]["conv_ms_per_step"], 3), d["clocks"]["reasons"])
    print("   train", json.dumps(d.get("train"))[:600])
except Exception as e:
    print("ws=${ws} failed", e); print(open("gpurun_out/${tag}_ws${ws}.err").read()[-1200:])
PY
done
paste -d'|' <(cut -d'|' -f1,2 gpurun_out/${tag}_conv_table_ws1.txt | head -8) <(cut -d'|' -f2 gpurun_out/${tag}_conv_table_ws2.txt | head -8)
