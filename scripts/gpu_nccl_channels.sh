#!/bin/bash
# N-GPU A/B of NCCL's channel count for the in-graph gradient reduction (training leg at B = 16 / GPU): the reduction needs
# ~30 GB/s, so a handful of channels leaves more SMs to the backward it overlaps with
n=${1:-4}
mkdir -p gpurun_out
for ch in default 4 2 default; do
  if [ "$ch" = "default" ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 --train-batches 16 --no-cpu-baseline --no-fp32 > gpurun_out/nccl_ch.json 2> gpurun_out/nccl_ch.err
  python - "$ch" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/nccl_ch.json").read().strip().splitlines()[-1])
    t = d["train"]["batch16"]
    print("NCCL_MAX_NCHANNELS=" + sys.argv[1], "train B=16/GPU:", round(t["ms_per_step"], 3), "ms", round(t["images_per_s"]), "img/s; inference", round(d["ms_per_step"], 3), "ms")
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/nccl_ch.err").read()[-600:])
PY
done
