#!/bin/bash
# A/B of environment settings on one GPU: bench (30 steps) per setting, each given as one quoted "VAR=val VAR2=val" argument
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 600 python -m pytest $TESTS -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -5; fi
for setting in "$@"; do
  env $setting TEDM_BENCH_CONV_TABLE=gpurun_out/abenv_table.txt timeout 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fp32 --no-train > gpurun_out/abenv.json 2> gpurun_out/abenv.err
  python - "$setting" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/abenv.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "->", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms conv", round(d["roofline"]["frac"], 4), d["clocks"]["reasons"])
    import os, re
    rows = open("gpurun_out/abenv_table.txt").read().splitlines()[1:]
    pat = os.environ.get("ROWS")
    for r in (rows[:1] if not pat else [r for r in rows if re.search(pat, r)]):
        print("   ", r)
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/abenv.err").read()[-800:])
PY
done
