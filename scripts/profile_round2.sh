#!/bin/bash
# Round-2 evidence (one gpurun call, 1 GPU): launch lists of the inference and training steps, ncu --set full of the conv,
# LinearAttention (tcgen05) and weight-gradient kernels, per-shape conv tables, the bench line.  Raw artefacts land in
# gpurun_out/; the text summaries are copied into profiles/ by hand (profiles/README.md indexes them).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
B="python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline --no-fp32"
# 1. the bench itself (never under a profiler), with the per-shape conv table
TEDM_BENCH_CONV_TABLE=gpurun_out/r02_infer_conv_table.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
# 2. launch lists (cold-cache, serialised: shares only)
timeout 500 ncu $M -c 2500 --log-file gpurun_out/r02_a_launches_infer.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline --no-fp32 > gpurun_out/r02_ncu_a.log 2>&1
timeout 500 ncu $M -c 3000 --log-file gpurun_out/r02_b_launches_train_b64.csv env GRAPH=0 WARM=1 STEPS=2 python scripts/train_bench.py 64 > gpurun_out/r02_ncu_b.log 2>&1
# 3. ncu --set full: every conv launch of one eager inference pass -> DRAM traffic of the dominant launch
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_igemm_kernel -s 62 -c 62 -o gpurun_out/r02_conv_full -f $B > gpurun_out/r02_ncu_conv.log 2>&1
ncu -i gpurun_out/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw.csv 2>/dev/null
python scripts/conv_traffic.py gpurun_out/r02_conv_full_raw.csv gpurun_out/r02_conv_traffic.json > /dev/null
# 4. ncu --set full: the tcgen05 LinearAttention kernels and the weight-gradient kernels
TEDM_PROF_TC_ONLY=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:linattn_tc_ -c 6 -o gpurun_out/r02_linattn_tc_full -f python scripts/prof_linattn.py 128 > gpurun_out/r02_ncu_la.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_wgrad3_kernel|conv_wgrad_kernel" -s 40 -c 12 -o gpurun_out/r02_wgrad_full -f env GRAPH=0 WARM=1 STEPS=1 python scripts/train_bench.py 64 > gpurun_out/r02_ncu_wg.log 2>&1
for r in r02_conv_full r02_linattn_tc_full r02_wgrad_full; do python scripts/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_summary.txt 2>/dev/null; done
# 5. per-shape conv table of the training step, LinearAttention timings, training step times
timeout 300 python scripts/train_conv_table.py 64 > gpurun_out/r02_train_conv_table_b64.txt 2>&1
timeout 200 python scripts/prof_linattn.py 128 > gpurun_out/r02_linattn_timing.txt 2>&1
for f in r02_a_launches_infer r02_b_launches_train_b64; do python scripts/summarize_launches.py gpurun_out/$f.csv > gpurun_out/$f.summary.txt 2>/dev/null; done
rm -f gpurun_out/r02_conv_full.ncu-rep   # 62 full-set launches: too large to bring back; the raw CSV and the summary stay
ls -la gpurun_out | grep r02_ | head -40
tail -c 300 gpurun_out/r02_bench_1gpu.err
