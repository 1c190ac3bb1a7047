#!/bin/bash
# Round-2 evidence refresh after the four-row weight-stationary conv kernel (one gpurun call, 1 GPU): the bench line, the
# inference launch list, ncu --set full of every conv launch of one inference pass (-> roofline.traffic), the training table.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
B="python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline --no-fp32"
TEDM_BENCH_CONV_TABLE=gpurun_out/r02_infer_conv_table.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 500 ncu $M -c 2500 --log-file gpurun_out/r02_a_launches_infer.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline --no-fp32 > gpurun_out/r02_ncu_a.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_igemm_kernel|conv_ws4_kernel" -s 62 -c 62 -o gpurun_out/r02_conv_full -f $B > gpurun_out/r02_ncu_conv.log 2>&1
ncu -i gpurun_out/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw.csv 2>/dev/null
python scripts/conv_traffic.py gpurun_out/r02_conv_full_raw.csv gpurun_out/r02_conv_traffic.json > /dev/null
python scripts/ncu_summary.py gpurun_out/r02_conv_full.ncu-rep > gpurun_out/r02_conv_full_summary.txt 2>/dev/null
# keep the source-level view of the new kernel only
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_ws4_kernel -s 12 -c 3 -o gpurun_out/r02_ws4_full -f $B > gpurun_out/r02_ncu_ws4.log 2>&1
rm -f gpurun_out/r02_conv_full.ncu-rep
if [ "$1" = "train" ]; then
  timeout 500 ncu $M -c 3000 --log-file gpurun_out/r02_b_launches_train_b64.csv env GRAPH=0 WARM=1 STEPS=2 python scripts/train_bench.py 64 > gpurun_out/r02_ncu_b.log 2>&1
  timeout 300 python scripts/train_conv_table.py 64 > gpurun_out/r02_train_conv_table_b64.txt 2>&1
fi
for f in r02_a_launches_infer r02_b_launches_train_b64; do python scripts/summarize_launches.py gpurun_out/$f.csv > gpurun_out/$f.summary.txt 2>/dev/null; done
ls -la gpurun_out | grep r02_ | head -40
tail -c 300 gpurun_out/r02_bench_1gpu.err
