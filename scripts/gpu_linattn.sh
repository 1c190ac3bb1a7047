#!/bin/bash
# quick iteration on the tcgen05 LinearAttention block: parity tests, timing against the mma.sync kernels, per-kernel ncu times
tag=${1:-la}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -k block_tc -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | tail -16
timeout 200 python scripts/prof_linattn.py 128 2>&1 | grep -v "^$"
TEDM_PROF_TC_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python scripts/prof_linattn.py 128 > gpurun_out/${tag}_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/${tag}_launches.csv 2>/dev/null | grep -i "linattn\|launches"
