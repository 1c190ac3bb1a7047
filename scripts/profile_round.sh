#!/bin/bash
# Round-end evidence: launch lists of the inference and training steps + ncu --set full of the kernels added this round.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
timeout 500 ncu $M -c 1500 --log-file gpurun_out/launches_infer3.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/ncu_infer3.log 2>&1
timeout 500 ncu $M -c 3000 --log-file gpurun_out/launches_train3.csv env GRAPH=0 WARM=1 STEPS=2 python scripts/train_bench.py 64 > gpurun_out/ncu_train3.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"linattn_fused|head_tail_mma|stem_conv7x7_mma|attention_flash|time_proj" -s 30 -c 14 -o gpurun_out/r01_new_kernels -f python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/ncu_new.log 2>&1
TEDM_BENCH_CONV_TABLE=gpurun_out/conv_table3.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/bench_i.log 2>&1
tail -c 300 gpurun_out/ncu_new.log
