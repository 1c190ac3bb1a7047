#!/bin/bash
# generic A/B on one GPU: parity tests, then the inference (+ training with TRAIN=1) bench with environment variable $1 set to $2 and to $3
var=$1; v1=$2; v2=$3; tag=${4:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_e2e.py tests/test_gpu_train.py tests/test_gpu_modules.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15
extra="--no-train"; [ "$TRAIN" = "1" ] && extra=""
for v in $v1 $v2; do
  env $var=$v TEDM_BENCH_CONV_TABLE=gpurun_out/${tag}_conv_table_${v}.txt timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 $extra > gpurun_out/${tag}_bench_${v}.json 2> gpurun_out/${tag}_${v}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${v}.json").read().strip().splitlines()[-1])
    print("$var=${v}", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), round(d["roofline"]["conv_ms_per_step"], 3), d["clocks"]["reasons"], "launches", d.get("gpu_launches"))
    t = d.get("train") or {}
    print("   train", {k: (round(x["ms_per_step"], 2), round(x["frac_of_sustained_bf16_peak"], 3)) for k, x in t.items() if isinstance(x, dict) and "ms_per_step" in x})
except Exception as e:
    print("$var=${v} failed", e); print(open("gpurun_out/${tag}_${v}.err").read()[-1500:])
PY
done
