#!/bin/bash
# One gpurun call: every GPU test file in its own process (a trapped kernel must not poison the rest), then the bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag] [extra pytest args]'
tag=${1:-check}
mkdir -p gpurun_out
out=gpurun_out/${tag}_tests.log
: > $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $out 2>&1
# the newest kernel first, on its own: if it is broken the rest of the run falls back to the kernels it replaces
if [ "${TC_FIRST:-1}" = "1" ]; then
  echo "=== tcgen05 LinearAttention block" >> $out
  if timeout 300 python -m pytest tests/test_gpu_ops.py -k block_tc -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/${tag}_tc.log 2>&1; then
    echo "tc block: PASS" >> $out
  else
    echo "tc block: FAIL -> TEDM_LINATTN_TC=0 for the rest" >> $out
    export TEDM_LINATTN_TC=0
  fi
  tail -25 gpurun_out/${tag}_tc.log >> $out
  timeout 200 python scripts/prof_linattn.py 128 >> $out 2>&1
fi
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> $out
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -60 >> $out
done
grep -E "^=== |passed|failed|error" $out | tail -40
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  tail -c 600 gpurun_out/${tag}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks") if k in d}
    keep["roofline"] = {k: d["roofline"].get(k) for k in ("achieved", "frac", "conv_ms_per_step", "conv_share_of_step")}
    keep["gn"] = {k: d["roofline_hbm"].get(k) for k in ("achieved", "frac", "ms_per_step")}
    keep["train"] = {k: {kk: v.get(kk) for kk in ("ms_per_step", "images_per_s", "tflops_fwd_bwd")} for k, v in d.get("train", {}).items() if isinstance(v, dict) and k.startswith("batch")}
    keep["cpu_baseline"] = d.get("cpu_baseline")
    print(json.dumps(keep, indent=1))
except Exception as e:
    print("bench parse failed:", e)
PY
fi
if [ "${NCU_LINATTN:-0}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_linattn_launches.csv python scripts/prof_linattn.py 128 > gpurun_out/${tag}_ncu_linattn.log 2>&1
  python scripts/summarize_launches.py gpurun_out/${tag}_linattn_launches.csv 2>/dev/null | head -30
fi
