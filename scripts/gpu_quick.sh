#!/bin/bash
# quick check of a conv change on one GPU: the conv / e2e parity tests, then the inference bench line with its conv table
tag=${1:-quick}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -6
TEDM_BENCH_CONV_TABLE=gpurun_out/${tag}_conv_table.txt timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), round(d["roofline"]["conv_ms_per_step"], 3), d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}.err").read()[-1500:])
PY
head -${2:-12} gpurun_out/${tag}_conv_table.txt
