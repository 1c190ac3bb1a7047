#!/bin/bash
# One GPU-box pass: parity tests per file (a faulting kernel poisons only its own process), smoke, short bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in probe conv ops e2e backward train head_train seg; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -s --tb=short -m gpu > gpurun_out/t_$t.log 2>&1
  echo "== test_gpu_$t exit $? =="; tail -n 4 gpurun_out/t_$t.log
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "== bench exit $? =="; tail -n 2 gpurun_out/bench.log
