#!/bin/bash
# conv A/B: CTA pairs on N <= 128 tiles only (TEDM_CTA_PAIRS=3) against every N (1): pair tests, bench lines, per-shape tables
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --no-header -p no:cacheprovider -k "pairs" 2>&1 | tail -5
for v in 1 3; do
  TEDM_CTA_PAIRS=$v TEDM_BENCH_CONV_TABLE=gpurun_out/r02s_conv_table_${v}.txt timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-train > gpurun_out/r02s_bench_${v}.json 2> gpurun_out/r02s_${v}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02s_bench_${v}.json").read().strip().splitlines()[-1])
print("pairs=${v}", round(d["value"], 1), round(d["ms_per_step"], 3), "conv", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), round(d["roofline"]["conv_ms_per_step"], 3), d["clocks"]["reasons"])
PY
done
paste -d'|' <(cut -d'|' -f1,2 gpurun_out/r02s_conv_table_3.txt | head -14) <(cut -d'|' -f2 gpurun_out/r02s_conv_table_1.txt | head -14)
