#!/bin/bash
# N-GPU data-parallel check: DP tests, then the bench (training leg at B = 16 / GPU) with and without the overlapped reduction
n=${1:-2}
tag=${2:-dp}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "^NCCL\|^$" | tail -15
for ov in 1 0; do
  TEDM_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --train-batches 16 64 --no-cpu-baseline --no-fp32 > gpurun_out/${tag}_bench_ov${ov}.json 2> gpurun_out/${tag}_bench_ov${ov}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_ov${ov}.json").read().strip().splitlines()[-1])
    print("overlap=${ov}", "infer", round(d["value"]), {k: (round(v["ms_per_step"], 3), round(v["images_per_s"])) for k, v in d["train"].items() if isinstance(v, dict) and "ms_per_step" in v})
except Exception as e:
    print("overlap=${ov} parse failed", e); print(open("gpurun_out/${tag}_bench_ov${ov}.err").read()[-1500:])
PY
done
