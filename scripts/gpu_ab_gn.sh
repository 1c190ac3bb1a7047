#!/bin/bash
# A/B of the GroupNorm-into-conv fusion on one GPU: off / everywhere / without the four-row kernel's variant / large tensors only
mkdir -p gpurun_out
run() {
  env "$@" timeout 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fp32 --no-train > gpurun_out/abgn.json 2> gpurun_out/abgn.err
  python - "$*" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/abgn.json").read().strip().splitlines()[-1])
print(sys.argv[1], "->", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["reasons"])
PY
}
run TEDM_FUSE_GN=0
run TEDM_FUSE_GN=1
run TEDM_FUSE_GN=1 TEDM_FUSE_GN_WS4=0
run TEDM_FUSE_GN=1 TEDM_FUSE_GN_WS4=0 TEDM_FUSE_GN_MIN_MB=100
run TEDM_FUSE_GN=1 TEDM_FUSE_GN_MIN_MB=100
run TEDM_FUSE_GN=0
