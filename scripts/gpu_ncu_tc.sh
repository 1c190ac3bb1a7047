#!/bin/bash
# ncu --set full of the tcgen05 LinearAttention kernels (one launch each), raw CSV pages brought back for reading here
tag=${1:-ncu_tc}
mkdir -p gpurun_out
TEDM_PROF_TC_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:linattn_tc_ -c 6 -o gpurun_out/${tag} -f python scripts/prof_linattn.py 128 > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
tail -3 gpurun_out/${tag}.log
ls -la gpurun_out/${tag}*
